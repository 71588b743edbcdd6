// AssessNet stem (models/assessment.py:54-57):  x = conv1(f) + conv1_p(p); bn1; relu; maxpool(3, 2, 1).
// The two 7x7 stride-2 convolutions are evaluated as ONE 4-input-channel convolution over the NHWC
// crop the ROI sampler wrote (channels 0..2 = normalised RGB, 3 = probability), with bn1 + ReLU fused
// in the epilogue; the 3x3/2 max-pool is a second, bandwidth-bound kernel.
//
// stem_conv_kernel: fp32 CUDA-core direct convolution.  One CTA = 8 x 16 output pixels x 64 channels;
// the 21 x 37 x 4 input patch and the 196 x 64 weight matrix live in shared memory.
// Algorithmic work: 2 * 196 * 64 * 128 * 128 = 0.411 GFLOP per (frame, object).
#include "ivosw_internal.h"

namespace ivosw {

constexpr int ST_TH = 8, ST_TW = 16;             // output tile
constexpr int ST_PH = ST_TH * 2 + 5, ST_PW = ST_TW * 2 + 5;  // 21 x 37 input patch
constexpr int ST_K = 196;                         // 7 * 7 * 4

__global__ void __launch_bounds__(256) stem_conv_kernel(const float4* __restrict__ crop,  // [B][256][256] float4
                                                        const float* __restrict__ wgt,    // [196][64]
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ shift,
                                                        float* __restrict__ c1,           // [B][128][128][64]
                                                        int raw) {                        // 1: convolution only (training)
    extern __shared__ float smem[];
    float* sw = smem;                       // 196 * 64
    float4* sp = reinterpret_cast<float4*>(smem + ST_K * 64);  // 21 * 37 float4
    const int b = blockIdx.z;
    const int oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;
    for (int i = threadIdx.x; i < ST_K * 64 / 4; i += 256)
        reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(wgt) + i);
    const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
    for (int i = threadIdx.x; i < ST_PH * ST_PW; i += 256) {
        int py = i / ST_PW, px = i - py * ST_PW;
        int iy = iy0 + py, ix = ix0 + px;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < ROI && ix >= 0 && ix < ROI) v = __ldg(crop + ((long long)b * ROI + iy) * ROI + ix);
        sp[i] = v;
    }
    __syncthreads();
    // thread -> 4 channels x 8 pixels (one column of the 8 x 16 tile)
    const int cg = threadIdx.x & 15;        // channel group: channels 4*cg .. 4*cg+3
    const int px = threadIdx.x >> 4;        // 0..15 tile column
    float acc[ST_TH][4];
#pragma unroll
    for (int r = 0; r < ST_TH; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    for (int kh = 0; kh < 7; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) {
            const float4* wrow = reinterpret_cast<const float4*>(sw + ((kh * 7 + kw) * 4) * 64) + cg;
            const float4 w0 = wrow[0], w1 = wrow[16], w2 = wrow[32], w3 = wrow[48];  // cin 0..3
#pragma unroll
            for (int r = 0; r < ST_TH; ++r) {
                const float4 v = sp[(r * 2 + kh) * ST_PW + px * 2 + kw];
                acc[r][0] = fmaf(v.x, w0.x, acc[r][0]); acc[r][1] = fmaf(v.x, w0.y, acc[r][1]);
                acc[r][2] = fmaf(v.x, w0.z, acc[r][2]); acc[r][3] = fmaf(v.x, w0.w, acc[r][3]);
                acc[r][0] = fmaf(v.y, w1.x, acc[r][0]); acc[r][1] = fmaf(v.y, w1.y, acc[r][1]);
                acc[r][2] = fmaf(v.y, w1.z, acc[r][2]); acc[r][3] = fmaf(v.y, w1.w, acc[r][3]);
                acc[r][0] = fmaf(v.z, w2.x, acc[r][0]); acc[r][1] = fmaf(v.z, w2.y, acc[r][1]);
                acc[r][2] = fmaf(v.z, w2.z, acc[r][2]); acc[r][3] = fmaf(v.z, w2.w, acc[r][3]);
                acc[r][0] = fmaf(v.w, w3.x, acc[r][0]); acc[r][1] = fmaf(v.w, w3.y, acc[r][1]);
                acc[r][2] = fmaf(v.w, w3.z, acc[r][2]); acc[r][3] = fmaf(v.w, w3.w, acc[r][3]);
            }
        }
    }
    const float4 sc = raw ? make_float4(1.f, 1.f, 1.f, 1.f) : __ldg(reinterpret_cast<const float4*>(scale) + cg);
    const float4 sh = raw ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4*>(shift) + cg);
    const float floor_ = raw ? -INFINITY : 0.f;
#pragma unroll
    for (int r = 0; r < ST_TH; ++r) {
        float4 o;
        o.x = fmaxf(fmaf(acc[r][0], sc.x, sh.x), floor_);
        o.y = fmaxf(fmaf(acc[r][1], sc.y, sh.y), floor_);
        o.z = fmaxf(fmaf(acc[r][2], sc.z, sh.z), floor_);
        o.w = fmaxf(fmaf(acc[r][3], sc.w, sh.w), floor_);
        long long pix = ((long long)b * 128 + (oy0 + r)) * 128 + (ox0 + px);
        reinterpret_cast<float4*>(c1 + pix * 64)[cg] = o;
    }
}

// maxpool 3x3 stride 2 pad 1 over NHWC [B][128][128][64] -> [B][64][64][64]; padding never wins
// (inputs are post-ReLU, but -inf semantics are kept for generality).
__global__ void __launch_bounds__(256) maxpool_kernel(const float4* __restrict__ c1, float4* __restrict__ out,
                                                      long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over B*64*64*16 float4
    if (i >= total) return;
    int cg = (int)(i & 15);
    long long p = i >> 4;
    int ox = (int)(p & 63), oy = (int)((p >> 6) & 63);
    long long b = p >> 12;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        int iy = oy * 2 + dy;
        if (iy < 0 || iy >= 128) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            int ix = ox * 2 + dx;
            if (ix < 0 || ix >= 128) continue;
            float4 v = __ldg(c1 + ((b * 128 + iy) * 128 + ix) * 16 + cg);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    out[i] = m;
}

static int stem_attr(ivosw_ctx* c, size_t smem) {
    static bool attr_set[64] = {};          // per device (the attribute is not process-wide)
    const int dv = c->device & 63;
    if (!attr_set[dv]) {
        IVOSW_CUDA(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dv] = true;
    }
    return IVOSW_OK;
}

// the 4-channel 7x7/2 convolution alone, on explicit operands (train.cu): crop [B][256][256][4], wgt [196][64]
int launch_stem_conv_raw(ivosw_ctx* c, const float* crop, const float* wgt_kc, float* out, int B, cudaStream_t s) {
    const size_t smem = (size_t)ST_K * 64 * 4 + (size_t)ST_PH * ST_PW * 16;
    int rc;
    if ((rc = stem_attr(c, smem))) return rc;
    dim3 grid(128 / ST_TW, 128 / ST_TH, B);
    stem_conv_kernel<<<grid, 256, smem, s>>>((const float4*)crop, wgt_kc, nullptr, nullptr, out, 1);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int launch_stem(ivosw_ctx* c, int B, cudaStream_t s) {
    const size_t smem = (size_t)ST_K * 64 * 4 + (size_t)ST_PH * ST_PW * 16;
    int rc;
    if ((rc = stem_attr(c, smem))) return rc;
    dim3 grid(128 / ST_TW, 128 / ST_TH, B);
    stem_conv_kernel<<<grid, 256, smem, s>>>((const float4*)c->crop.p, c->stem_w, c->stem_scale, c->stem_shift,
                                             (float*)c->c1.p, 0);
    IVOSW_CUDA(cudaGetLastError());
    long long total = (long long)B * 64 * 64 * 16;
    maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const float4*)c->c1.p, (float4*)c->pool.p, total);
    IVOSW_CUDA(cudaGetLastError());
    c->launches += 2;
    return IVOSW_OK;
}

}  // namespace ivosw
