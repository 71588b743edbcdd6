// In-repo arithmetic of the ATNet round wrapper utils/utils_atnet.py::run_VOS_singleiact (the networks themselves,
// forward_ANet / forward_TNet / encoder_3ch, are external and stay external):
//
//   :95-96    ReflectionPad2d(pad_info[1] + pad_info[0]) of the n_obj x 3 x H x W scribble planes
//   :101-102, 124-126  sigmoid of the logits; prob_onehot_t = prob[:, 0]
//   :146-150  alpha-blend against the previous round's probabilities, prob_map_of_frames[frame] = result
//   :157-159  all_P = cat([zeros, prob_map_of_frames], 1)[:, :, hpad1:-hpad2, wpad1:-wpad2]
//
// All three are HBM-bound element-wise passes; algorithmic bytes: pad  n_obj*3*(H*W read + PH*PW written)*4,
// blend  n_obj*PH*PW*(2 reads + 2 writes)*4 per frame,  assemble  T*(O*H*W read + (O+1)*H*W written)*4.
// The reference materialises sigmoid, two scaled temporaries and their sum per frame, and at the end a zeros tensor,
// the concatenation (T*(O+1)*PH*PW) and a strided slice view of it; here each result is written once.
#include "ivosw_internal.h"

namespace ivosw {

// out[n][c][y][x] = in[n][c][reflect(y - top)][reflect(x - left)]   (torch.nn.ReflectionPad2d: no edge repeat)
__global__ void __launch_bounds__(256) reflect_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w,
                                                          int left, int top, int PH, int PW) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const long long plane = blockIdx.z;
    if (x >= PW) return;
    int sy = y - top, sx = x - left;
    sy = sy < 0 ? -sy : (sy >= h ? 2 * (h - 1) - sy : sy);
    sx = sx < 0 ? -sx : (sx >= w ? 2 * (w - 1) - sx : sx);
    out[plane * PH * PW + (long long)y * PW + x] = __ldg(in + plane * h * w + (long long)sy * w + sx);
}

// prob = sigmoid(logit);  blended = alpha * prob + beta * prev  (two rounded products, one rounded sum: torch's
// (alpha * p) + ((1 - alpha) * prev) evaluates exactly that way); blended may alias prev (in-place update of
// prob_map_of_frames[frame]).  has_prev == 0: blended = prob (the annotated frame, :105-106 + :150).
__global__ void __launch_bounds__(256) sigmoid_blend_kernel(const float* __restrict__ logit, const float* prev,
                                                            float* __restrict__ prob, float* blended, long long n,
                                                            float alpha, float beta, int has_prev) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 3 < n && (((uintptr_t)(logit + i) | (uintptr_t)(prob + i) | (uintptr_t)(blended + i) |
                       (has_prev ? (uintptr_t)(prev + i) : 0)) & 15) == 0) {
        const float4 x = *reinterpret_cast<const float4*>(logit + i);
        float4 p;
        p.x = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x.x))); p.y = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x.y)));
        p.z = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x.z))); p.w = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x.w)));
        *reinterpret_cast<float4*>(prob + i) = p;
        float4 o = p;
        if (has_prev) {
            const float4 q = *reinterpret_cast<const float4*>(prev + i);
            o.x = __fadd_rn(__fmul_rn(alpha, p.x), __fmul_rn(beta, q.x));
            o.y = __fadd_rn(__fmul_rn(alpha, p.y), __fmul_rn(beta, q.y));
            o.z = __fadd_rn(__fmul_rn(alpha, p.z), __fmul_rn(beta, q.z));
            o.w = __fadd_rn(__fmul_rn(alpha, p.w), __fmul_rn(beta, q.w));
        }
        *reinterpret_cast<float4*>(blended + i) = o;
        return;
    }
    for (long long j = i; j < n && j < i + 4; ++j) {
        const float p = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-logit[j])));
        prob[j] = p;
        blended[j] = has_prev ? __fadd_rn(__fmul_rn(alpha, p), __fmul_rn(beta, prev[j])) : p;
    }
}

// all_P[t][0] = 0, all_P[t][o + 1][y][x] = prob_map[t][o][y0 + y][x0 + x]
__global__ void __launch_bounds__(256) atnet_assemble_kernel(const float* __restrict__ prob_map, float* __restrict__ all_p,
                                                             int O, int PH, int PW, int y0, int x0, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int t = blockIdx.z / (O + 1), ch = blockIdx.z % (O + 1);
    if (x >= W) return;
    float v = 0.f;
    if (ch > 0) v = __ldg(prob_map + (((long long)t * O + (ch - 1)) * PH + (y0 + y)) * PW + (x0 + x));
    all_p[(((long long)t * (O + 1) + ch) * H + y) * W + x] = v;
}

int launch_reflect_pad(ivosw_ctx* c, const float* in, float* out, long long planes, int h, int w, int left, int right, int top,
                       int bottom, cudaStream_t s) {
    const int PH = h + top + bottom, PW = w + left + right;
    dim3 grid((PW + 255) / 256, PH, (unsigned)planes);
    reflect_pad_kernel<<<grid, 256, 0, s>>>(in, out, h, w, left, top, PH, PW);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int launch_sigmoid_blend(ivosw_ctx* c, const float* logit, const float* prev, float* prob, float* blended, long long n,
                         float alpha, float beta, cudaStream_t s) {
    const long long thr = (n + 3) / 4;
    sigmoid_blend_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, s>>>(logit, prev, prob, blended, n, alpha, beta,
                                                                     prev != nullptr ? 1 : 0);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int launch_atnet_assemble(ivosw_ctx* c, const float* prob_map, float* all_p, int T, int O, int PH, int PW, int y0, int x0,
                          int H, int W, cudaStream_t s) {
    dim3 grid((W + 255) / 256, H, (unsigned)(T * (O + 1)));
    atnet_assemble_kernel<<<grid, 256, 0, s>>>(prob_map, all_p, O, PH, PW, y0, x0, H, W);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

}  // namespace ivosw
