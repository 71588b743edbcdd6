// Tail of the MANet round wrapper utils/utils_manet.py::get_results: for every frame the (O+1)-channel
// logits at embedding resolution h x w are upsampled bilinearly (align_corners=True) to H x W
// (:76-77, 109-110, 146-147), the per-pixel argmax gives the mask (:78-79, 113-114, 148-149) and the
// channel softmax of the stacked upsampled logits gives all_P (:161).  The reference materialises T
// separate H x W logit tensors, concatenates them and launches softmax on the 315 MB result; here one
// kernel reads the small logits (L2-resident) and writes masks and all_P once.
//
// HBM-bound: algorithmic bytes = T*(C+1)*H*W*4 written (+ T*C*h*w*4 read).
#include "ivosw_internal.h"

namespace ivosw {

constexpr int TAIL_MAXC = 16;

__global__ void __launch_bounds__(256) manet_tail_kernel(const float* __restrict__ logits, int C, int h, int w, int H,
                                                         int W, float sy, float sx, float* __restrict__ masks,
                                                         float* __restrict__ all_p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int t = blockIdx.z;
    if (x >= W) return;
    // ATen area_pixel_compute_source_index(align_corners=True): src = scale * dst, fp32
    const float fy = __fmul_rn(sy, (float)y), fx = __fmul_rn(sx, (float)x);
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly1 = fminf(fmaxf(__fsub_rn(fy, (float)y0), 0.f), 1.f), ly0 = __fsub_rn(1.f, ly1);
    const float lx1 = fminf(fmaxf(__fsub_rn(fx, (float)x0), 0.f), 1.f), lx0 = __fsub_rn(1.f, lx1);
    const float* base = logits + (long long)t * C * h * w;
    float v[TAIL_MAXC];
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int c = 0; c < TAIL_MAXC; ++c) {
        if (c < C) {
            const float* p = base + (long long)c * h * w;
            const float v00 = __ldg(p + y0 * w + x0), v01 = __ldg(p + y0 * w + x1);
            const float v10 = __ldg(p + y1 * w + x0), v11 = __ldg(p + y1 * w + x1);
            const float r0 = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
            const float r1 = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
            v[c] = __fadd_rn(__fmul_rn(ly0, r0), __fmul_rn(ly1, r1));
            if (v[c] > best) { best = v[c]; bi = c; }   // first maximum wins (torch.argmax)
        }
    }
    const long long pix = (long long)y * W + x;
    const long long HW = (long long)H * W;
    if (masks) masks[(long long)t * HW + pix] = (float)bi;
    if (all_p) {
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < TAIL_MAXC; ++c)
            if (c < C) { v[c] = expf(v[c] - best); sum += v[c]; }
#pragma unroll
        for (int c = 0; c < TAIL_MAXC; ++c)
            if (c < C) all_p[((long long)t * C + c) * HW + pix] = __fdiv_rn(v[c], sum);
    }
}

int launch_manet_tail(ivosw_ctx* c, const float* logits, int T, int C, int h, int w, int H, int W, float* masks,
                      float* all_p, cudaStream_t s) {
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    dim3 grid((W + 255) / 256, H, T);
    manet_tail_kernel<<<grid, 256, 0, s>>>(logits, C, h, w, H, W, sy, sx, masks, all_p);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

// ---------------------------------------------------------------------------------------------------------
// utils/utils_manet.py::rough_ROI (22-39): first-round scribble labels are kept only inside the bounding box
// (+-20 px) of the pixels that are not -1; everything outside becomes 0.  Slice ends are exclusive and clamp to
// h-1 / w-1 exactly as the reference's Python slices do (SURVEY A.Q12).  One CTA per image.
__global__ void __launch_bounds__(256) rough_roi_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w,
                                                        int dist, int* __restrict__ empty_flag) {
    __shared__ int s_mn[2], s_mx[2];
    const int b = blockIdx.x;
    const float* src = in + (long long)b * h * w;
    float* dst = out + (long long)b * h * w;
    if (threadIdx.x == 0) { s_mn[0] = s_mn[1] = 0x7fffffff; s_mx[0] = s_mx[1] = -1; }
    __syncthreads();
    int ymin = 0x7fffffff, xmin = 0x7fffffff, ymax = -1, xmax = -1;
    for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
        if (src[i] != -1.0f) {
            const int y = i / w, x = i - y * w;
            ymin = min(ymin, y); ymax = max(ymax, y); xmin = min(xmin, x); xmax = max(xmax, x);
        }
    }
    atomicMin(&s_mn[0], ymin); atomicMin(&s_mn[1], xmin); atomicMax(&s_mx[0], ymax); atomicMax(&s_mx[1], xmax);
    __syncthreads();
    if (s_mx[0] < 0) {      // torch.min over an empty nonzero() raises in the reference
        if (threadIdx.x == 0) atomicExch(empty_flag, 1);
        return;
    }
    const int y0 = max(s_mn[0] - dist, 0), y1 = min(s_mx[0] + dist, h - 1);   // [y0, y1) rows kept
    const int x0 = max(s_mn[1] - dist, 0), x1 = min(s_mx[1] + dist, w - 1);
    for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
        const int y = i / w, x = i - y * w;
        dst[i] = (y >= y0 && y < y1 && x >= x0 && x < x1) ? src[i] : 0.0f;
    }
}

int launch_rough_roi(ivosw_ctx* c, const float* in, float* out, int B, int h, int w, int dist, int* empty_flag_dev,
                     cudaStream_t s) {
    IVOSW_CUDA(cudaMemsetAsync(empty_flag_dev, 0, sizeof(int), s));
    rough_roi_kernel<<<B, 256, 0, s>>>(in, out, h, w, dist, empty_flag_dev);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

}  // namespace ivosw
