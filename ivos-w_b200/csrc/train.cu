// One optimisation step of the quality-assessment network on the device: quality_assessment.py::train, loop body
// :240-269 (BASELINE config C5).
//
//   :213      assess_net.train()        BatchNorm with BATCH statistics, running statistics updated (momentum 0.1, unbiased
//                                       variance), in all 54 BatchNorm layers
//   :240      forward                   bbox -> ROI crop (roi.cu, as in inference) -> 4-channel stem -> maxpool -> 16
//                                       bottlenecks -> mean -> fc1
//   :251-262  loss                      mean over the samples with union > 0 of (pred - target)^2
//   :265      backward                  NO zero_grad anywhere in the loop: gradients ACCUMULATE from step to step
//   :266-268  clamp_(-1, 1) in place    the clamped value is what keeps accumulating
//   :269      SGD(lr, momentum, wd)     d = g + wd * p;  buf = momentum * buf + d;  p -= lr * buf
//
// Arithmetic: activations, parameters, gradients and optimiser state are fp32 (NHWC); where the FLOPs are —
//   forward conv / dgrad   the bottleneck convolutions run on conv_tc.cu's tcgen05 kernel in its fp32-grade split-fp16 mode
//                          (split x and w into hi / lo fp16 planes, convolve with an identity BatchNorm, merge; data gradients
//                          are pre-scaled by a power of two taken from their largest magnitude so that 1e-8-sized values stay
//                          inside fp16's normal range).  A data gradient is itself a convolution of dy with the transposed,
//                          tap-flipped weight (stride 2: of the zero-upsampled dy; 1x1 stride 2: at low resolution, then
//                          scattered to the even pixels).  conv_simt.cu's fp32 CUDA-core kernel remains for shapes the GEMM
//                          tiling does not cover (B * OH * OW not a multiple of 128) and as IVOSW_TRAIN_TC=0 for validation;
//                          the 4-channel stem is the fp32 direct convolution of stem.cu.
//   wgrad                  dW[co][tap][ci] = sum_m dy[m][co] x[m @ tap][ci]: 64 x 64 tiles, the M axis split over CTAs into
//                          partial sums, reduced in a fixed order (deterministic, no atomics)
//   BatchNorm              per-channel sums in double, fixed-order reduction; backward with the batch-statistics terms
// Layout: activations NHWC fp32; parameters, gradients and momentum are flat arrays in the order of ivosw_assess_load's
// blob, so an exported blob loads straight back into the inference path.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ivosw_internal.h"

namespace ivosw {

constexpr int TC_BM_ROWS = 128;      // pixels per GEMM tile of conv_tc.cu

// conv_simt.cu
int launch_conv_simt_raw(ivosw_ctx* c, const float* in, const float* wgt, const float* scale, const float* shift,
                         const float* residual, float* out, int B, int in_hw, int cin, int out_hw, int cout, int k, int stride,
                         int pad, int relu, cudaStream_t s);
// stem.cu
int launch_stem_conv_raw(ivosw_ctx* c, const float* crop, const float* wgt_kc, float* out, int B, cudaStream_t s);

namespace {

constexpr float BN_MOMENTUM = 0.1f;

struct TrainLayer {                 // one convolution + its BatchNorm (index 0 = the stem)
    int cin, cout, k, stride, pad, in_hw, out_hw;
    size_t w_off, g_off, b_off, rm_off, rv_off;     // offsets into the blob (floats)
    // activations (floats from the arena base)
    const float* x = nullptr;       // input activation
    float* y = nullptr;             // raw convolution output
    float* a = nullptr;             // after BatchNorm (+ residual) (+ ReLU)
    float* mean = nullptr;          // batch statistics
    float* invstd = nullptr;
    float* wT = nullptr;            // transposed / flipped weight for the data gradient
};

struct TrainState {
    int cap = 0;                                    // units the arena is sized for
    size_t n_blob = 0;
    float* blob = nullptr;                          // parameters + BatchNorm buffers, blob order
    float* gnew = nullptr;                          // gradient of the current step
    float* gacc = nullptr;                          // accumulated, clamped gradient (what .grad holds in the reference)
    float* mom = nullptr;                           // SGD momentum buffers
    std::vector<TrainLayer> L;                      // [0] stem, [1..52] bottleneck convolutions
    size_t fc_off = 0;
    DeviceBuffer arena, scratch, wt_arena, stats, ws;
    DeviceBuffer tc_in, tc_out, tc_w;               // split-fp16 operand planes of the tensor-core convolutions
    float* tc_scale = nullptr;                      // [2]: power-of-two pre-scale of a data gradient and its inverse
    float* identity_scale = nullptr;                // 2048 ones / zeros for raw convolutions
    float* identity_shift = nullptr;
    unsigned char* pool_idx = nullptr;
    float *crop = nullptr, *pool = nullptr, *gap = nullptr, *pred = nullptr, *dpred = nullptr, *loss_dev = nullptr;
    float* stem_wkc = nullptr;                      // stem weight transposed to [196][64] for stem_conv_kernel
    long long steps = 0;
};

// ---------------------------------------------------------------------------------------------- BatchNorm
// partial sums over a row chunk: part[(split * C + c) * 2 + {0, 1}] = sum y, sum y^2 (double)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ y, long long M, int C, int rows_per_split,
                                                       double* __restrict__ part) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int lane_row = threadIdx.x >> 5;                       // 8 row lanes
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = min(M, r0 + rows_per_split);
    double s = 0.0, q = 0.0;
    long long r = r0 + lane_row;
    for (; r + 24 < r1; r += 32) {                    // four rows in flight per thread; same summation order as the tail loop
        const float v0 = y[r * C + c], v1 = y[(r + 8) * C + c], v2 = y[(r + 16) * C + c], v3 = y[(r + 24) * C + c];
        s += v0; q += (double)v0 * v0; s += v1; q += (double)v1 * v1;
        s += v2; q += (double)v2 * v2; s += v3; q += (double)v3 * v3;
    }
    for (; r < r1; r += 8) {
        const float v = y[r * C + c];
        s += v; q += (double)v * v;
    }
    __shared__ double sh[2][8][32];
    sh[0][lane_row][threadIdx.x & 31] = s; sh[1][lane_row][threadIdx.x & 31] = q;
    __syncthreads();
    if (lane_row == 0) {
        for (int i = 1; i < 8; ++i) { s += sh[0][i][threadIdx.x & 31]; q += sh[1][i][threadIdx.x & 31]; }
        part[((size_t)blockIdx.y * C + c) * 2 + 0] = s;
        part[((size_t)blockIdx.y * C + c) * 2 + 1] = q;
    }
}

// mean / invstd of the batch; running statistics (momentum 0.1, unbiased variance) written into the blob
// (block = 32 channels x 8 split lanes: lane j adds splits j, j + 8, ...; the eight partial sums are combined in a fixed order)
__device__ __forceinline__ void sum_splits_8(const double* __restrict__ part, int splits, int C, int c, double& a, double& b) {
    __shared__ double sh[2][8][32];
    const int lane_row = threadIdx.x >> 5, cl = threadIdx.x & 31;
    double s = 0.0, q = 0.0;
    if (c < C)
        for (int i = lane_row; i < splits; i += 8) { s += part[((size_t)i * C + c) * 2]; q += part[((size_t)i * C + c) * 2 + 1]; }
    sh[0][lane_row][cl] = s; sh[1][lane_row][cl] = q;
    __syncthreads();
    a = 0.0; b = 0.0;
    if (lane_row == 0)
        for (int i = 0; i < 8; ++i) { a += sh[0][i][cl]; b += sh[1][i][cl]; }
}

__global__ void __launch_bounds__(256) bn_finalize_kernel(const double* __restrict__ part, int splits, int C, long long M,
                                                          float* __restrict__ mean, float* __restrict__ invstd,
                                                          float* __restrict__ run_mean, float* __restrict__ run_var) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    double s, q;
    sum_splits_8(part, splits, C, c, s, q);
    if ((threadIdx.x >> 5) != 0 || c >= C) return;
    const double m = s / (double)M;
    double var = q / (double)M - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)BN_EPS));
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    run_mean[c] = (float)((1.0 - BN_MOMENTUM) * (double)run_mean[c] + BN_MOMENTUM * m);
    run_var[c] = (float)((1.0 - BN_MOMENTUM) * (double)run_var[c] + BN_MOMENTUM * unbiased);
}

// a = gamma * (y - mean) * invstd + beta (+ residual) (ReLU)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ residual,
                                                       float* __restrict__ a, long long total4, int C, int relu) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c = (int)((i * 4) % C);
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    const float4 m = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = (v.x - m.x) * is.x * g.x + b.x; o.y = (v.y - m.y) * is.y * g.y + b.y;
    o.z = (v.z - m.z) * is.z * g.z + b.z; o.w = (v.w - m.w) * is.w * g.w + b.w;
    if (residual) {
        const float4 r = reinterpret_cast<const float4*>(residual)[i];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    reinterpret_cast<float4*>(a)[i] = o;
}

// backward, pass 1: d_out masked by the ReLU (act > 0) -> partial sums of dz and dz * xhat (double)
__global__ void __launch_bounds__(256) bn_bwd_stats_kernel(const float* __restrict__ dout, const float* __restrict__ act,
                                                           const float* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, long long M, int C,
                                                           int rows_per_split, double* __restrict__ part) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int lane_row = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = min(M, r0 + rows_per_split);
    const float m = mean[c], is = invstd[c];
    double sb = 0.0, sg = 0.0;
    long long r = r0 + lane_row;
    for (; r + 24 < r1; r += 32) {                    // four rows (twelve loads) in flight per thread, same summation order
        float dz[4], yv[4], av[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long o = (r + 8 * u) * C + c;
            dz[u] = dout[o]; yv[u] = y[o]; av[u] = act ? act[o] : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float d = (av[u] > 0.f) ? dz[u] : 0.f;
            sb += d; sg += (double)d * (double)((yv[u] - m) * is);
        }
    }
    for (; r < r1; r += 8) {
        float dz = dout[r * C + c];
        if (act && !(act[r * C + c] > 0.f)) dz = 0.f;
        sb += dz; sg += (double)dz * (double)((y[r * C + c] - m) * is);
    }
    __shared__ double sh[2][8][32];
    sh[0][lane_row][threadIdx.x & 31] = sb; sh[1][lane_row][threadIdx.x & 31] = sg;
    __syncthreads();
    if (lane_row == 0) {
        for (int i = 1; i < 8; ++i) { sb += sh[0][i][threadIdx.x & 31]; sg += sh[1][i][threadIdx.x & 31]; }
        part[((size_t)blockIdx.y * C + c) * 2 + 0] = sb;
        part[((size_t)blockIdx.y * C + c) * 2 + 1] = sg;
    }
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const double* __restrict__ part, int splits, int C,
                                                              float* __restrict__ dbeta, float* __restrict__ dgamma,
                                                              float* __restrict__ sums /*[2][C]*/) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    double sb, sg;
    sum_splits_8(part, splits, C, c, sb, sg);
    if ((threadIdx.x >> 5) != 0 || c >= C) return;
    dbeta[c] = (float)sb; dgamma[c] = (float)sg;
    sums[c] = (float)sb; sums[C + c] = (float)sg;
}

// backward, pass 2: dy = gamma * invstd * (dz - dbeta / M - xhat * dgamma / M); dz_out (optional) receives the masked dz
// (the gradient that also flows into the identity branch of a residual sum)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ act,
                                                           const float* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ sums, float inv_m, float* __restrict__ dy,
                                                           float* __restrict__ dz_out, long long total4, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // four consecutive channels of one pixel
    if (i >= total4) return;
    const int c = (int)((i * 4) % C);
    const float4 d4 = reinterpret_cast<const float4*>(dout)[i], y4 = reinterpret_cast<const float4*>(y)[i];
    float dz[4] = {d4.x, d4.y, d4.z, d4.w};
    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
    if (act) {
        const float4 a4 = reinterpret_cast<const float4*>(act)[i];
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) if (!(av[u] > 0.f)) dz[u] = 0.f;
    }
    if (dz_out) reinterpret_cast<float4*>(dz_out)[i] = make_float4(dz[0], dz[1], dz[2], dz[3]);
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float xh = (yv[u] - mean[c + u]) * invstd[c + u];
        o[u] = gamma[c + u] * invstd[c + u] * (dz[u] - sums[c + u] * inv_m - xh * sums[C + c + u] * inv_m);
    }
    reinterpret_cast<float4*>(dy)[i] = make_float4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------- pooling / head
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                          unsigned char* __restrict__ idx, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B*64*64*64
    if (i >= total) return;
    const int c = (int)(i & 63);
    const long long p = i >> 6;
    const int ox = (int)(p & 63), oy = (int)((p >> 6) & 63);
    const long long b = p >> 12;
    float m = -INFINITY; int best = 0;
    for (int dy = 0; dy < 3; ++dy) {
        const int iy = oy * 2 - 1 + dy;
        if (iy < 0 || iy >= 128) continue;
        for (int dx = 0; dx < 3; ++dx) {
            const int ix = ox * 2 - 1 + dx;
            if (ix < 0 || ix >= 128) continue;
            const float v = in[((b * 128 + iy) * 128 + ix) * 64 + c];
            if (v > m) { m = v; best = dy * 3 + dx; }           // first maximum wins (ATen scans in this order)
        }
    }
    out[i] = m; idx[i] = (unsigned char)best;
}

// gradient of the max-pool by GATHER: every input pixel collects from the (at most 4) windows whose arg-max it is
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dout, const unsigned char* __restrict__ idx,
                                                          float* __restrict__ din, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B*128*128*64
    if (i >= total) return;
    const int c = (int)(i & 63);
    const long long p = i >> 6;
    const int ix = (int)(p & 127), iy = (int)((p >> 7) & 127);
    const long long b = p >> 14;
    float g = 0.f;
    for (int oy = (iy >> 1); oy <= ((iy + 1) >> 1); ++oy) {
        if (oy < 0 || oy >= 64) continue;
        const int dy = iy - (oy * 2 - 1);
        if (dy < 0 || dy > 2) continue;
        for (int ox = (ix >> 1); ox <= ((ix + 1) >> 1); ++ox) {
            if (ox < 0 || ox >= 64) continue;
            const int dx = ix - (ox * 2 - 1);
            if (dx < 0 || dx > 2) continue;
            const long long o = ((b * 64 + oy) * 64 + ox) * 64 + c;
            if (idx[o] == dy * 3 + dx) g += dout[o];
        }
    }
    din[i] = g;
}

// gap[b][c] = mean over the 64 pixels; pred[b] = fc(gap)
__global__ void __launch_bounds__(256) gap_fc_fwd_kernel(const float* __restrict__ r5, const float* __restrict__ fcw,
                                                         const float* __restrict__ fcb, float* __restrict__ gap,
                                                         float* __restrict__ pred) {
    const int b = blockIdx.x;
    const float* x = r5 + (long long)b * 64 * 2048;
    float part = 0.f;
    for (int ch = threadIdx.x; ch < 2048; ch += 256) {
        float s = 0.f;
        for (int p = 0; p < 64; ++p) s += x[p * 2048 + ch];
        const float g = s * (1.0f / 64.0f);
        gap[(long long)b * 2048 + ch] = g;
        part = fmaf(g, fcw[ch], part);
    }
    __shared__ float red[8];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        pred[b] = t + fcb[0];
    }
}

// masked MSE (:251-262) and its gradient: loss = sum_valid (pred - t)^2 / counter; dpred = 2 (pred - t) / counter
__global__ void loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, const int* __restrict__ valid, int B,
                            float* __restrict__ dpred, float* __restrict__ loss_out /*[2]: loss, counter*/) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int cnt = 0;
    for (int n = 0; n < B; ++n) cnt += valid[n] != 0;
    float loss = 0.f;
    for (int n = 0; n < B; ++n) {
        const float d = pred[n] - target[n];
        if (valid[n]) loss += d * d;                                  // F.mse_loss of a single element
        dpred[n] = (valid[n] && cnt > 0) ? 2.0f * d / (float)cnt : 0.f;
    }
    loss_out[0] = cnt > 0 ? loss / (float)cnt : 0.f;
    loss_out[1] = (float)cnt;
}

// fc backward: dW[c] = sum_b dpred[b] gap[b][c]; db = sum dpred; d r5[b][p][c] = dpred[b] * W[c] / 64
__global__ void __launch_bounds__(256) fc_bwd_kernel(const float* __restrict__ dpred, const float* __restrict__ gap,
                                                     const float* __restrict__ fcw, int B, float* __restrict__ dW,
                                                     float* __restrict__ db, float* __restrict__ dr5) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;           // 2048 channels
    if (ch < 2048) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s = fmaf(dpred[b], gap[(long long)b * 2048 + ch], s);
        dW[ch] = s;
        const float w = fcw[ch] * (1.0f / 64.0f);
        for (int b = 0; b < B; ++b) {
            const float g = dpred[b] * w;
            for (int p = 0; p < 64; ++p) dr5[((long long)b * 64 + p) * 2048 + ch] = g;
        }
    }
    if (ch == 0) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += dpred[b];
        db[0] = s;
    }
}

// ---------------------------------------------------------------------------------------------- convolution gradients
// wT[ci][k-1-kh][k-1-kw][co] = w[co][kh][kw][ci]: the weight of the convolution that computes the data gradient
__global__ void weight_transpose_kernel(const float* __restrict__ w, float* __restrict__ wT, int cout, int cin, int k) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)cout * k * k * cin;
    if (i >= total) return;
    const int ci = (int)(i % cin);
    long long t = i / cin;
    const int kw = (int)(t % k); t /= k;
    const int kh = (int)(t % k);
    const int co = (int)(t / k);
    wT[(((long long)ci * k + (k - 1 - kh)) * k + (k - 1 - kw)) * cout + co] = w[i];
}

// up[b][2oy][2ox][:] = dy[b][oy][ox][:], zeros elsewhere (stride-2 3x3 data gradient as a stride-1 convolution)
__global__ void zero_upsample_kernel(const float* __restrict__ dy, float* __restrict__ up, int ohw, int C, long long total4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B * (2 ohw)^2 * C / 4
    if (i >= total4) return;
    const int c4 = C / 4;
    const int c = (int)(i % c4);
    long long p = i / c4;
    const int iw = 2 * ohw;
    const int x = (int)(p % iw); p /= iw;
    const int y = (int)(p % iw);
    const long long b = p / iw;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(x & 1) && !(y & 1)) v = reinterpret_cast<const float4*>(dy)[((b * ohw + (y >> 1)) * ohw + (x >> 1)) * c4 + c];
    reinterpret_cast<float4*>(up)[i] = v;
}

// dx[b][2oy][2ox][:] (+)= low[b][oy][ox][:] (1x1 stride-2 data gradient computed at low resolution)
__global__ void scatter_add_stride2_kernel(const float* __restrict__ low, float* __restrict__ dx, int ohw, int C, long long total4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B * ohw^2 * C / 4
    if (i >= total4) return;
    const int c4 = C / 4;
    const int c = (int)(i % c4);
    long long p = i / c4;
    const int x = (int)(p % ohw); p /= ohw;
    const int y = (int)(p % ohw);
    const long long b = p / ohw;
    const int iw = 2 * ohw;
    float4* d = reinterpret_cast<float4*>(dx) + ((b * iw + 2 * y) * iw + 2 * x) * c4 + c;
    const float4 v = reinterpret_cast<const float4*>(low)[i];
    float4 o = *d;
    o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
    *d = o;
}

__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, long long total4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    float4 x = reinterpret_cast<float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    reinterpret_cast<float4*>(a)[i] = x;
}

// Weight gradient.  Flattened k = tap * cin + ci (the blob's OHWI order); one CTA = 64 output channels x 64 k values
// over rows [m0, m1) of the M = B * OH * OW axis, as a partial sum part[split][co][k].
struct WgradGeom { int B, H, W, Cin, OH, OW, Cout, k, stride, pad, K; long long M; int rows_per_split; };

__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, WgradGeom g,
                                                    float* __restrict__ part) {
    __shared__ __align__(16) float Ds[16][64 + 4];      // dy rows x 64 output channels
    __shared__ __align__(16) float Xs[16][64 + 4];      // the same rows' inputs x 64 k values
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * 64, co0 = blockIdx.y * 64;
    const long long m0 = (long long)blockIdx.z * g.rows_per_split, m1 = min(g.M, m0 + g.rows_per_split);
    // loader: thread -> row (tid / 16) of the 16-row slab, 4 consecutive columns at (tid % 16) * 4
    const int lr = tid >> 4, lc = (tid & 15) * 4;
    const int kk = k0 + lc;                             // this thread's first k (4 consecutive ci inside one tap: Cin % 4 == 0)
    const bool k_ok = kk < g.K;
    const int tap = k_ok ? kk / g.Cin : 0, ci = k_ok ? kk - tap * g.Cin : 0;
    const int kh = tap / g.k, kw = tap - kh * g.k;
    const int ty = tid >> 4, tx = tid & 15;             // 16 x 16 threads, each 4 co x 4 k
    float acc[4][4] = {};
    // one 16-row slab ahead: the next slab's global loads are issued before this slab's 256 FMAs per thread (the arithmetic
    // and its order are unchanged)
    auto load_slab = [&](long long mb, float4& dv, float4& xv) {
        const long long m = mb + lr;
        dv = make_float4(0.f, 0.f, 0.f, 0.f); xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < m1) {
            dv = __ldg(reinterpret_cast<const float4*>(dy + m * g.Cout + co0 + lc));
            if (k_ok) {
                const int ow = (int)(m % g.OW);
                const long long t = m / g.OW;
                const int oh = (int)(t % g.OH);
                const long long b = t / g.OH;
                const int ih = oh * g.stride - g.pad + kh, iw = ow * g.stride - g.pad + kw;
                if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                    xv = __ldg(reinterpret_cast<const float4*>(x + ((b * g.H + ih) * g.W + iw) * g.Cin + ci));
            }
        }
    };
    float4 dv, xv;
    if (m0 < m1) load_slab(m0, dv, xv);
    for (long long mb = m0; mb < m1; mb += 16) {
        __syncthreads();
        *reinterpret_cast<float4*>(&Ds[lr][lc]) = dv;
        *reinterpret_cast<float4*>(&Xs[lr][lc]) = xv;
        __syncthreads();
        if (mb + 16 < m1) load_slab(mb + 16, dv, xv);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float4 d4 = *reinterpret_cast<const float4*>(&Ds[r][ty * 4]);
            const float4 x4 = *reinterpret_cast<const float4*>(&Xs[r][tx * 4]);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w}, xx[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d[i], xx[j], acc[i][j]);
        }
    }
    float* out = part + (size_t)blockIdx.z * g.Cout * g.K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < g.K) out[(size_t)co * g.K + k] = acc[i][j];
        }
    }
}

// The same weight gradient on 128 x 128 tiles (layers with at least 128 output channels and 128 k values): every thread owns
// 8 output channels x 8 k values, so a shared-memory row feeds 64 FMAs per four 16-byte reads (16 per two in the 64 x 64
// kernel).  Row order and per-output summation order are those of wgrad_kernel.
__global__ void __launch_bounds__(256) wgrad128_kernel(const float* __restrict__ dy, const float* __restrict__ x, WgradGeom g,
                                                       float* __restrict__ part) {
    __shared__ __align__(16) float Ds[16][128 + 4];
    __shared__ __align__(16) float Xs[16][128 + 4];
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * 128, co0 = blockIdx.y * 128;
    const long long m0 = (long long)blockIdx.z * g.rows_per_split, m1 = min(g.M, m0 + g.rows_per_split);
    // loader: thread -> row (tid / 16) of the 16-row slab, two float4 at columns (tid % 16) * 4 and 64 + (tid % 16) * 4
    const int lr = tid >> 4, lc = (tid & 15) * 4;
    int tapv[2], civ[2], khv[2], kwv[2]; bool okv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int kk = k0 + h * 64 + lc;
        okv[h] = kk < g.K;
        tapv[h] = okv[h] ? kk / g.Cin : 0; civ[h] = okv[h] ? kk - tapv[h] * g.Cin : 0;
        khv[h] = tapv[h] / g.k; kwv[h] = tapv[h] - khv[h] * g.k;
    }
    const int ty = tid >> 4, tx = tid & 15;             // 16 x 16 threads: co = {ty*4.., 64+ty*4..}, k = {tx*4.., 64+tx*4..}
    float acc[8][8] = {};
    auto load_slab = [&](long long mb, float4* dv, float4* xv) {
        const long long m = mb + lr;
        dv[0] = dv[1] = xv[0] = xv[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < m1) {
            dv[0] = __ldg(reinterpret_cast<const float4*>(dy + m * g.Cout + co0 + lc));
            dv[1] = __ldg(reinterpret_cast<const float4*>(dy + m * g.Cout + co0 + 64 + lc));
            const int ow = (int)(m % g.OW);
            const long long t = m / g.OW;
            const int oh = (int)(t % g.OH);
            const long long b = t / g.OH;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!okv[h]) continue;
                const int ih = oh * g.stride - g.pad + khv[h], iw = ow * g.stride - g.pad + kwv[h];
                if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                    xv[h] = __ldg(reinterpret_cast<const float4*>(x + ((b * g.H + ih) * g.W + iw) * g.Cin + civ[h]));
            }
        }
    };
    float4 dv[2], xv[2];
    if (m0 < m1) load_slab(m0, dv, xv);
    for (long long mb = m0; mb < m1; mb += 16) {
        __syncthreads();
        *reinterpret_cast<float4*>(&Ds[lr][lc]) = dv[0]; *reinterpret_cast<float4*>(&Ds[lr][64 + lc]) = dv[1];
        *reinterpret_cast<float4*>(&Xs[lr][lc]) = xv[0]; *reinterpret_cast<float4*>(&Xs[lr][64 + lc]) = xv[1];
        __syncthreads();
        if (mb + 16 < m1) load_slab(mb + 16, dv, xv);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float4 da = *reinterpret_cast<const float4*>(&Ds[r][ty * 4]), db = *reinterpret_cast<const float4*>(&Ds[r][64 + ty * 4]);
            const float4 xa = *reinterpret_cast<const float4*>(&Xs[r][tx * 4]), xb = *reinterpret_cast<const float4*>(&Xs[r][64 + tx * 4]);
            const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
            const float xx[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(d[i], xx[j], acc[i][j]);
        }
    }
    float* out = part + (size_t)blockIdx.z * g.Cout * g.K;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int co = co0 + (i >> 2) * 64 + ty * 4 + (i & 3);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + (j >> 2) * 64 + tx * 4 + (j & 3);
            if (k < g.K) out[(size_t)co * g.K + k] = acc[i][j];
        }
    }
}

// grad[i] = sum over splits (fixed order)
__global__ void reduce_splits_kernel(const float* __restrict__ part, int splits, long long n, float* __restrict__ grad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];
    grad[i] = s;
}

// :265-269 for one contiguous range of trainable parameters
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ gnew, float* __restrict__ gacc, float* __restrict__ mom,
                           long long n, float lr, float momentum, float wd, int apply) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = gacc[i] + gnew[i];                        // backward() accumulates into .grad (never zeroed)
    g = fminf(fmaxf(g, -1.0f), 1.0f);                   // clamp_(-1, 1), in place
    gacc[i] = g;
    if (!apply) return;
    const float d = g + wd * p[i];
    const float b = momentum * mom[i] + d;              // first step: buf = d (momentum buffer starts at zero)
    mom[i] = b;
    p[i] = p[i] - lr * b;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host
// ---------------------------------------------------------------------------------------------- tensor-core convolutions
// Forward convolutions and data gradients of the bottleneck layers run on conv_tc.cu's tcgen05 kernel (split-fp16 operands,
// fp32 accumulation: the fp32-grade arithmetic of the inference path).  The training step keeps fp32 NHWC activations, so
// a convolution is: split x and w into (hi, lo) fp16 planes, conv_tc with an identity BatchNorm, merge the planes back.
// Data gradients are tiny (1e-3 .. 1e-8, far below fp16's normal range): they are multiplied by a power of two chosen from
// their largest magnitude while splitting and divided by it while merging — exact, and entirely on the device.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n4, unsigned int* __restrict__ out) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        m = fmaxf(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))), m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));     // non-negative floats order like their bits
}
// scale[0] = 2^e with max * 2^e in [512, 1024), scale[1] = 2^-e (1, 1 for an all-zero tensor); resets the max
__global__ void pick_scale_kernel(unsigned int* __restrict__ maxbits, float* __restrict__ scale) {
    const float m = __uint_as_float(*maxbits);
    int e = 0;
    if (m > 0.f && isfinite(m)) { int ex; frexpf(m, &ex); e = 10 - ex; }      // m = f * 2^ex, f in [0.5, 1)
    e = max(-60, min(60, e));
    scale[0] = ldexpf(1.f, e); scale[1] = ldexpf(1.f, -e);
    *maxbits = 0u;
}
__global__ void __launch_bounds__(256) split_scaled_kernel(const float* __restrict__ in, __half* __restrict__ hi,
                                                           __half* __restrict__ lo, long long n2, const float* __restrict__ scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const float sc = scale ? scale[0] : 1.f;
    float2 v = reinterpret_cast<const float2*>(in)[i];
    v.x = fminf(fmaxf(v.x * sc, -65504.f), 65504.f); v.y = fminf(fmaxf(v.y * sc, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(v.x, v.y);
    const float2 hf = __half22float2(h);
    reinterpret_cast<__half2*>(hi)[i] = h;
    reinterpret_cast<__half2*>(lo)[i] = __floats2half2_rn((v.x - hf.x) * 2048.0f, (v.y - hf.y) * 2048.0f);
}
__global__ void __launch_bounds__(256) merge_scaled_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                                           float* __restrict__ out, long long n2, const float* __restrict__ scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const float inv = scale ? scale[1] : 1.f;
    const float2 h = __half22float2(reinterpret_cast<const __half2*>(hi)[i]);
    const float2 l = __half22float2(reinterpret_cast<const __half2*>(lo)[i]);
    reinterpret_cast<float2*>(out)[i] = make_float2(fmaf(l.x, 1.0f / 2048.0f, h.x) * inv, fmaf(l.y, 1.0f / 2048.0f, h.y) * inv);
}

static bool train_tc_enabled() {                 // read per call: tests compare the two arithmetic paths in one process
    const char* e = getenv("IVOSW_TRAIN_TC");
    return !(e && atoi(e) == 0);
}
static bool tc_eligible(int B, int cin, int out_hw, int cout) {
    return train_tc_enabled() && cin % 64 == 0 && cout % 64 == 0 && ((long long)B * out_hw * out_hw) % TC_BM_ROWS == 0;
}
// y[B][out_hw][out_hw][cout] = conv(x[B][in_hw][in_hw][cin], w[cout][k][k][cin]); dyn_scale: x is a data gradient
static int conv_f32_tc(ivosw_ctx* c, TrainState* T, const float* x, const float* w, float* y, int B, int in_hw, int cin, int out_hw,
                       int cout, int k, int stride, int pad, bool dyn_scale, cudaStream_t s) {
    const long long n_in = (long long)B * in_hw * in_hw * cin, n_out = (long long)B * out_hw * out_hw * cout;
    const long long n_w = (long long)cout * k * k * cin;
    const SplitAct in = split_view(T->tc_in), out = split_view(T->tc_out), ws = split_view(T->tc_w);
    const float* sc = nullptr;
    if (dyn_scale) {
        unsigned int* mb = reinterpret_cast<unsigned int*>(T->tc_scale + 2);
        absmax_kernel<<<592, 256, 0, s>>>(x, n_in / 4, mb);
        pick_scale_kernel<<<1, 1, 0, s>>>(mb, T->tc_scale);
        c->launches += 2;
        sc = T->tc_scale;
    }
    split_scaled_kernel<<<(unsigned)((n_in / 2 + 255) / 256), 256, 0, s>>>(x, in.hi, in.lo, n_in / 2, sc);
    split_scaled_kernel<<<(unsigned)((n_w / 2 + 255) / 256), 256, 0, s>>>(w, ws.hi, ws.lo, n_w / 2, nullptr);
    c->launches += 2;
    IVOSW_CUDA(cudaGetLastError());
    ConvLayer L{};
    L.cin = cin; L.cout = cout; L.k = k; L.stride = stride; L.pad = pad; L.in_hw = in_hw; L.out_hw = out_hw;
    L.relu = false; L.residual = 0; L.is_downsample = false; L.first_of_block = false;
    L.scale = T->identity_scale; L.shift = T->identity_shift; L.w_hi = ws.hi; L.w_lo = ws.lo;
    int rc;
    if ((rc = launch_conv_tc(c, L, in, nullptr, out, B, 3, s))) return rc;
    merge_scaled_kernel<<<(unsigned)((n_out / 2 + 255) / 256), 256, 0, s>>>(out.hi, out.lo, y, n_out / 2, sc);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

static int n4(long long n) { return (int)((n / 4 + 255) / 256); }

static int bn_forward(ivosw_ctx* c, TrainState* T, TrainLayer& L, long long M, const float* residual, int relu, cudaStream_t s) {
    const int C = L.cout;
    const int splits = (int)std::min<long long>(256, std::max<long long>(1, M / 512));
    const int rows = (int)((M + splits - 1) / splits);
    double* part = (double*)T->stats.p;
    bn_stats_kernel<<<dim3(C / 32, splits), 256, 0, s>>>(L.y, M, C, rows, part);
    bn_finalize_kernel<<<(C + 31) / 32, 256, 0, s>>>(part, splits, C, M, L.mean, L.invstd, T->blob + L.rm_off, T->blob + L.rv_off);
    const long long tot4 = M * C / 4;
    bn_apply_kernel<<<(unsigned)((tot4 + 255) / 256), 256, 0, s>>>(L.y, L.mean, L.invstd, T->blob + L.g_off, T->blob + L.b_off,
                                                                  residual, L.a, tot4, C, relu);
    c->launches += 3;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

// d_out (gradient w.r.t. L.a) -> dy (gradient w.r.t. L.y), dgamma / dbeta into gnew; relu_act: mask by L.a > 0
static int bn_backward(ivosw_ctx* c, TrainState* T, TrainLayer& L, long long M, const float* dout, bool relu_masked, float* dy,
                       float* dz_out, cudaStream_t s) {
    const int C = L.cout;
    const int splits = (int)std::min<long long>(256, std::max<long long>(1, M / 512));
    const int rows = (int)((M + splits - 1) / splits);
    double* part = (double*)T->stats.p;
    float* sums = (float*)((char*)T->stats.p + sizeof(double) * 2 * 256 * 2048);
    const float* act = relu_masked ? L.a : nullptr;
    bn_bwd_stats_kernel<<<dim3(C / 32, splits), 256, 0, s>>>(dout, act, L.y, L.mean, L.invstd, M, C, rows, part);
    bn_bwd_finalize_kernel<<<(C + 31) / 32, 256, 0, s>>>(part, splits, C, T->gnew + L.b_off, T->gnew + L.g_off, sums);
    const long long tot4 = M * C / 4;
    bn_bwd_apply_kernel<<<(unsigned)((tot4 + 255) / 256), 256, 0, s>>>(dout, act, L.y, L.mean, L.invstd, T->blob + L.g_off, sums,
                                                                      1.0f / (float)M, dy, dz_out, tot4, C);
    c->launches += 3;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

static int conv_wgrad(ivosw_ctx* c, TrainState* T, const TrainLayer& L, int B, const float* dy, const float* x, cudaStream_t s) {
    WgradGeom g;
    g.B = B; g.H = L.in_hw; g.W = L.in_hw; g.Cin = L.cin; g.OH = L.out_hw; g.OW = L.out_hw; g.Cout = L.cout;
    g.k = L.k; g.stride = L.stride; g.pad = L.pad; g.K = L.k * L.k * L.cin;
    g.M = (long long)B * L.out_hw * L.out_hw;
    const bool wide = L.cout % 128 == 0 && g.K >= 128 && !(getenv("IVOSW_WGRAD_WIDE") && atoi(getenv("IVOSW_WGRAD_WIDE")) == 0);
    const int tile = wide ? 128 : 64;
    const int tiles = ((g.K + tile - 1) / tile) * (L.cout / tile);
    long long splits = std::max<long long>(1, std::min<long long>(1184 / tiles + 1, g.M / 256));
    if (splits > 512) splits = 512;
    g.rows_per_split = (int)(((g.M + splits - 1) / splits + 15) / 16 * 16);
    splits = (g.M + g.rows_per_split - 1) / g.rows_per_split;
    const size_t need = sizeof(float) * (size_t)splits * L.cout * g.K;
    int rc;
    if ((rc = ensure(T->ws, need))) return rc;
    if (wide) wgrad128_kernel<<<dim3((g.K + 127) / 128, L.cout / 128, (unsigned)splits), 256, 0, s>>>(dy, x, g, (float*)T->ws.p);
    else wgrad_kernel<<<dim3((g.K + 63) / 64, L.cout / 64, (unsigned)splits), 256, 0, s>>>(dy, x, g, (float*)T->ws.p);
    const long long n = (long long)L.cout * g.K;
    reduce_splits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const float*)T->ws.p, (int)splits, n, T->gnew + L.w_off);
    c->launches += 2;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

// dx (size B x in_hw^2 x cin) = data gradient of layer L for dy; `accumulate`: dx += instead of =
static int conv_dgrad(ivosw_ctx* c, TrainState* T, const TrainLayer& L, int B, const float* dy, float* dx, float* tmp_up,
                      float* tmp_low, bool accumulate, cudaStream_t s) {
    int rc;
    const long long wn = (long long)L.cout * L.k * L.k * L.cin;
    weight_transpose_kernel<<<(unsigned)((wn + 255) / 256), 256, 0, s>>>(T->blob + L.w_off, L.wT, L.cout, L.cin, L.k);
    c->launches += 1;
    const long long n_in = (long long)B * L.in_hw * L.in_hw * L.cin;
    if (L.stride == 1) {
        float* dst = accumulate ? tmp_low : dx;
        if (tc_eligible(B, L.cout, L.in_hw, L.cin)) {
            if ((rc = conv_f32_tc(c, T, dy, L.wT, dst, B, L.out_hw, L.cout, L.in_hw, L.cin, L.k, 1, L.k - 1 - L.pad, true, s))) return rc;
        } else if ((rc = launch_conv_simt_raw(c, dy, L.wT, T->identity_scale, T->identity_shift, nullptr, dst, B, L.out_hw, L.cout, L.in_hw,
                                       L.cin, L.k, 1, L.k - 1 - L.pad, 0, s)))
            return rc;
        if (accumulate) { add_inplace_kernel<<<n4(n_in), 256, 0, s>>>(dx, tmp_low, n_in / 4); c->launches += 1; }
    } else if (L.k == 3) {          // 3x3 stride 2: stride-1 convolution of the zero-upsampled dy with the flipped weight
        const long long up4 = (long long)B * L.in_hw * L.in_hw * L.cout / 4;
        zero_upsample_kernel<<<(unsigned)((up4 + 255) / 256), 256, 0, s>>>(dy, tmp_up, L.out_hw, L.cout, up4);
        c->launches += 1;
        float* dst = accumulate ? tmp_low : dx;
        if (tc_eligible(B, L.cout, L.in_hw, L.cin)) {
            if ((rc = conv_f32_tc(c, T, tmp_up, L.wT, dst, B, L.in_hw, L.cout, L.in_hw, L.cin, 3, 1, 1, true, s))) return rc;
        } else if ((rc = launch_conv_simt_raw(c, tmp_up, L.wT, T->identity_scale, T->identity_shift, nullptr, dst, B, L.in_hw, L.cout,
                                       L.in_hw, L.cin, 3, 1, 1, 0, s)))
            return rc;
        if (accumulate) { add_inplace_kernel<<<n4(n_in), 256, 0, s>>>(dx, tmp_low, n_in / 4); c->launches += 1; }
    } else {                        // 1x1 stride 2: at low resolution, then scattered to the even pixels
        if (tc_eligible(B, L.cout, L.out_hw, L.cin)) {
            if ((rc = conv_f32_tc(c, T, dy, L.wT, tmp_low, B, L.out_hw, L.cout, L.out_hw, L.cin, 1, 1, 0, true, s))) return rc;
        } else if ((rc = launch_conv_simt_raw(c, dy, L.wT, T->identity_scale, T->identity_shift, nullptr, tmp_low, B, L.out_hw, L.cout,
                                       L.out_hw, L.cin, 1, 1, 0, 0, s)))
            return rc;
        if (!accumulate) IVOSW_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * n_in, s));
        const long long low4 = (long long)B * L.out_hw * L.out_hw * L.cin / 4;
        scatter_add_stride2_kernel<<<(unsigned)((low4 + 255) / 256), 256, 0, s>>>(tmp_low, dx, L.out_hw, L.cin, low4);
        c->launches += 1;
    }
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int train_apply(ivosw_ctx* c, float lr, float momentum, float wd, cudaStream_t s);

void train_release(ivosw_ctx* c) {
    TrainState* T = static_cast<TrainState*>(c->train_state);
    if (!T) return;
    float* shifted[] = {T->blob, T->gnew, T->gacc, T->mom};
    for (float* p : shifted) if (p) cudaFree(p - 2);
    float* ptrs[] = {T->identity_scale, T->identity_shift, T->stem_wkc, T->tc_scale};
    for (float* p : ptrs) if (p) cudaFree(p);
    release(T->arena); release(T->scratch); release(T->wt_arena); release(T->stats); release(T->ws);
    release(T->tc_in); release(T->tc_out); release(T->tc_w);
    delete T;
    c->train_state = nullptr;
}

// (Re)starts training from `blob` (ivosw_assess_load's layout): parameters and BatchNorm buffers are copied to the device,
// gradients and momentum buffers start at zero.
int train_begin(ivosw_ctx* c, const float* blob, size_t n_floats) {
    train_release(c);
    TrainState* T = new TrainState();
    c->train_state = T;
    T->n_blob = n_floats;
    const size_t nb = sizeof(float) * n_floats;
    // The blob starts with 6 scalars (mean, std): with the arrays shifted by 2 floats every weight tensor and every
    // BatchNorm vector starts on a 16-byte boundary (all tensor sizes are multiples of 4), as the float4 loads need.
    float** arrs[] = {&T->blob, &T->gnew, &T->gacc, &T->mom};
    for (float** a : arrs) {
        float* raw = nullptr;
        IVOSW_CUDA(cudaMalloc(&raw, nb + 16));
        *a = raw + 2;
    }
    IVOSW_CUDA(cudaMemcpy(T->blob, blob, nb, cudaMemcpyHostToDevice));
    IVOSW_CUDA(cudaMemset(T->gnew, 0, nb)); IVOSW_CUDA(cudaMemset(T->gacc, 0, nb)); IVOSW_CUDA(cudaMemset(T->mom, 0, nb));
    {
        std::vector<float> ones(2048, 1.0f);
        IVOSW_CUDA(cudaMalloc(&T->identity_scale, 2048 * sizeof(float)));
        IVOSW_CUDA(cudaMalloc(&T->identity_shift, 2048 * sizeof(float)));
        IVOSW_CUDA(cudaMemcpy(T->identity_scale, ones.data(), 2048 * sizeof(float), cudaMemcpyHostToDevice));
        IVOSW_CUDA(cudaMemset(T->identity_shift, 0, 2048 * sizeof(float)));
        IVOSW_CUDA(cudaMalloc(&T->stem_wkc, 196 * 64 * sizeof(float)));
    }
    // layer table with blob offsets
    size_t off = 6;
    TrainLayer st{};
    st.cin = 4; st.cout = 64; st.k = 7; st.stride = 2; st.pad = 3; st.in_hw = 256; st.out_hw = 128;
    st.w_off = off; off += (size_t)64 * 196;
    st.g_off = off; st.b_off = off + 64; st.rm_off = off + 128; st.rv_off = off + 192; off += 256;
    T->L.push_back(st);
    for (const ConvLayer& Lh : c->layers) {
        TrainLayer L{};
        L.cin = Lh.cin; L.cout = Lh.cout; L.k = Lh.k; L.stride = Lh.stride; L.pad = Lh.pad; L.in_hw = Lh.in_hw; L.out_hw = Lh.out_hw;
        L.w_off = off; off += (size_t)Lh.cout * Lh.k * Lh.k * Lh.cin;
        L.g_off = off; L.b_off = off + Lh.cout; L.rm_off = off + 2 * (size_t)Lh.cout; L.rv_off = off + 3 * (size_t)Lh.cout;
        off += 4 * (size_t)Lh.cout;
        T->L.push_back(L);
    }
    T->fc_off = off;
    if (off + 2049 != n_floats) { set_error("AssessNet blob length"); return IVOSW_ERR_INVALID; }
    IVOSW_CUDA(cudaDeviceSynchronize());
    return IVOSW_OK;
}

static int train_workspace(ivosw_ctx* c, TrainState* T, int B) {
    if (B <= T->cap) return IVOSW_OK;
    int rc;
    // arena: per layer y and a; crop, pool, gap, pred ...; all offsets in floats
    size_t total = 0;
    auto take = [&](size_t n) { size_t o = total; total += (n + 63) & ~(size_t)63; return o; };
    std::vector<size_t> oy(T->L.size()), oa(T->L.size()), om(T->L.size()), oi(T->L.size()), ow(T->L.size());
    const size_t o_crop = take((size_t)B * 256 * 256 * 4), o_pool = take((size_t)B * 64 * 64 * 64);
    for (size_t i = 0; i < T->L.size(); ++i) {
        const TrainLayer& L = T->L[i];
        const size_t n = (size_t)B * L.out_hw * L.out_hw * L.cout;
        oy[i] = take(n); oa[i] = take(n); om[i] = take(L.cout); oi[i] = take(L.cout);
    }
    const size_t o_gap = take((size_t)B * 2048), o_pred = take(B), o_dpred = take(B), o_loss = take(64);
    const size_t o_idx = take(((size_t)B * 64 * 64 * 64 + 3) / 4);
    if ((rc = ensure(T->arena, total * sizeof(float)))) return rc;
    float* base = (float*)T->arena.p;
    T->crop = base + o_crop; T->pool = base + o_pool; T->gap = base + o_gap; T->pred = base + o_pred; T->dpred = base + o_dpred;
    T->loss_dev = base + o_loss; T->pool_idx = (unsigned char*)(base + o_idx);
    size_t wtot = 0;
    for (size_t i = 0; i < T->L.size(); ++i) {
        TrainLayer& L = T->L[i];
        L.y = base + oy[i]; L.a = base + oa[i]; L.mean = base + om[i]; L.invstd = base + oi[i];
        ow[i] = wtot; wtot += ((size_t)L.cout * L.k * L.k * L.cin + 63) & ~(size_t)63;
    }
    if ((rc = ensure(T->wt_arena, wtot * sizeof(float)))) return rc;
    for (size_t i = 0; i < T->L.size(); ++i) T->L[i].wT = (float*)T->wt_arena.p + ow[i];
    // gradient scratch: 5 buffers of the largest activation (c1: 128 x 128 x 64 = res2 output 64 x 64 x 256)
    const size_t big = (size_t)B * 128 * 128 * 64;
    if ((rc = ensure(T->scratch, 5 * big * sizeof(float)))) return rc;
    if ((rc = ensure(T->stats, sizeof(double) * 2 * 256 * 2048 + sizeof(float) * 2 * 2048))) return rc;
    // split-fp16 planes of the tensor-core convolutions: the largest activation (B x 64 x 64 x 256) and the largest weight
    if ((rc = ensure(T->tc_in, sizeof(float) * (size_t)B * 64 * 64 * 256))) return rc;
    if ((rc = ensure(T->tc_out, sizeof(float) * (size_t)B * 64 * 64 * 256))) return rc;
    if ((rc = ensure(T->tc_w, sizeof(float) * (size_t)512 * 3 * 3 * 512))) return rc;
    if (!T->tc_scale) {
        IVOSW_CUDA(cudaMalloc(&T->tc_scale, sizeof(float) * 4));
        IVOSW_CUDA(cudaMemset(T->tc_scale, 0, sizeof(float) * 4));
    }
    // wire the inputs: block structure as in capi.cu
    const float* x = T->pool;
    const float *t1 = nullptr, *t2 = nullptr;
    for (size_t i = 1; i < T->L.size(); ++i) {
        const ConvLayer& Lh = c->layers[i - 1];
        TrainLayer& L = T->L[i];
        if (Lh.first_of_block) { L.x = x; t1 = L.a; }
        else if (Lh.k == 3) { L.x = t1; t2 = L.a; }
        else if (Lh.is_downsample) { L.x = x; }
        else { L.x = t2; x = L.a; }
    }
    T->L[0].x = T->crop;
    T->cap = B;
    return IVOSW_OK;
}

int train_step(ivosw_ctx* c, const UnitAddr& ua, int B, int H, int W, const float* targets_dev, const int* valid_dev, float lr,
               float momentum, float wd, int apply, float* loss_host, float* pred_host, cudaStream_t s) {
    TrainState* T = static_cast<TrainState*>(c->train_state);
    if (!T) { set_error("ivosw_assess_train_begin has not been called"); return IVOSW_ERR_STATE; }
    int rc;
    if ((rc = train_workspace(c, T, B))) return rc;
    float* G = (float*)T->scratch.p;
    const size_t big = (size_t)T->cap * 128 * 128 * 64;
    float *gA = G, *gB = G + big, *gY = G + 2 * big, *gUp = G + 3 * big, *gLow = G + 4 * big;
    // ---------------------------------------------------------------- forward (:240)
    if ((rc = launch_bbox(c, ua, B, H, W, s))) return rc;
    {   // ROI crop into the training arena (fp32 NHWC, normalised RGB + probability)
        DeviceBuffer saved = c->crop;
        c->crop.p = T->crop; c->crop.bytes = (size_t)B * ROI * ROI * 16;
        rc = launch_roi_sample(c, ua, B, H, W, (float*)c->boxes.p, false, s);
        c->crop = saved;
        if (rc) return rc;
    }
    {   // stem: raw 4-channel 7x7/2 convolution, BatchNorm (batch statistics), ReLU, max-pool with arg-max
        TrainLayer& L = T->L[0];
        // stem weight [64][7][7][4] -> [(kh, kw, ci)][64] for the direct convolution kernel: a plain transpose
        weight_transpose_kernel<<<(64 * 196 + 255) / 256, 256, 0, s>>>(T->blob + L.w_off, gA, 64, 196, 1);
        IVOSW_CUDA(cudaMemcpyAsync(T->stem_wkc, gA, sizeof(float) * 196 * 64, cudaMemcpyDeviceToDevice, s));
        if ((rc = launch_stem_conv_raw(c, T->crop, T->stem_wkc, L.y, B, s))) return rc;
        if ((rc = bn_forward(c, T, L, (long long)B * 128 * 128, nullptr, 1, s))) return rc;
        const long long tot = (long long)B * 64 * 64 * 64;
        maxpool_fwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(L.a, T->pool, T->pool_idx, tot);
        c->launches += 2;
    }
    for (size_t i = 1; i < T->L.size(); ++i) {
        const ConvLayer& Lh = c->layers[i - 1];
        TrainLayer& L = T->L[i];
        if (tc_eligible(B, L.cin, L.out_hw, L.cout)) {
            if ((rc = conv_f32_tc(c, T, L.x, T->blob + L.w_off, L.y, B, L.in_hw, L.cin, L.out_hw, L.cout, L.k, L.stride, L.pad, false, s)))
                return rc;
        } else if ((rc = launch_conv_simt_raw(c, L.x, T->blob + L.w_off, T->identity_scale, T->identity_shift, nullptr, L.y, B, L.in_hw,
                                       L.cin, L.out_hw, L.cout, L.k, L.stride, L.pad, 0, s)))
            return rc;
        const long long M = (long long)B * L.out_hw * L.out_hw;
        const float* res = nullptr;
        int relu = 1;
        if (Lh.is_downsample) relu = 0;
        else if (!Lh.first_of_block && Lh.k == 1) {       // conv3: + identity (block input or the downsample branch's output)
            if (Lh.residual == 2) res = T->L[i - 1].a;    // the downsample layer precedes conv3 in the list
            else {
                // the block input: the activation conv1 of this block read
                res = T->L[i - 2].x;
            }
        }
        if ((rc = bn_forward(c, T, L, M, res, relu, s))) return rc;
    }
    const TrainLayer& last = T->L.back();
    gap_fc_fwd_kernel<<<B, 256, 0, s>>>(last.a, T->blob + T->fc_off, T->blob + T->fc_off + 2048, T->gap, T->pred);
    loss_kernel<<<1, 32, 0, s>>>(T->pred, targets_dev, valid_dev, B, T->dpred, T->loss_dev);
    c->launches += 2;
    IVOSW_CUDA(cudaGetLastError());
    if ((rc = ensure_pinned(c, sizeof(float) * (B + 8)))) return rc;
    float* pin = (float*)c->pinned_small;
    IVOSW_CUDA(cudaMemcpyAsync(pin, T->loss_dev, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    IVOSW_CUDA(cudaMemcpyAsync(pin + 8, T->pred, sizeof(float) * B, cudaMemcpyDeviceToHost, s));
    IVOSW_CUDA(cudaStreamSynchronize(s));
    if (pred_host) memcpy(pred_host, pin + 8, sizeof(float) * B);
    if (loss_host) *loss_host = pin[0];
    if (pin[1] == 0.f) {            // `if counter == 0: continue` (:259): no backward, no optimiser step
        if (loss_host) *loss_host = NAN;
        return IVOSW_OK;
    }
    // ---------------------------------------------------------------- backward (:265)
    IVOSW_CUDA(cudaMemsetAsync(T->gnew, 0, sizeof(float) * T->n_blob, s));
    fc_bwd_kernel<<<(2048 + 255) / 256, 256, 0, s>>>(T->dpred, T->gap, T->blob + T->fc_off, B, T->gnew + T->fc_off,
                                                     T->gnew + T->fc_off + 2048, gA);
    c->launches += 1;
    // walk the bottlenecks backwards; gA holds the gradient w.r.t. the current block's output
    int i = (int)T->L.size() - 1;
    while (i >= 1) {
        // block = [conv1, conv2, (downsample), conv3] ending at i
        const bool has_ds = c->layers[i - 2].is_downsample;      // layer list index of T->L[i - 1]
        const int i3 = i, id = has_ds ? i - 1 : -1, i2 = has_ds ? i - 2 : i - 1, i1 = i2 - 1;
        TrainLayer &L3 = T->L[i3], &L2 = T->L[i2], &L1 = T->L[i1];
        const long long M3 = (long long)B * L3.out_hw * L3.out_hw;
        const long long M1 = (long long)B * L1.out_hw * L1.out_hw;
        // conv3 + bn3 (+ identity) + relu: dS = dZ * 1[Z > 0] -> gB; dy3 -> gY
        if ((rc = bn_backward(c, T, L3, M3, gA, true, gY, gB, s))) return rc;
        if ((rc = conv_wgrad(c, T, L3, B, gY, L3.x, s))) return rc;
        if ((rc = conv_dgrad(c, T, L3, B, gY, gA, gUp, gLow, false, s))) return rc;          // gA <- d t2
        // conv2
        if ((rc = bn_backward(c, T, L2, M3, gA, true, gY, nullptr, s))) return rc;
        if ((rc = conv_wgrad(c, T, L2, B, gY, L2.x, s))) return rc;
        if ((rc = conv_dgrad(c, T, L2, B, gY, gA, gUp, gLow, false, s))) return rc;          // gA <- d t1
        // conv1
        if ((rc = bn_backward(c, T, L1, M1, gA, true, gY, nullptr, s))) return rc;
        if ((rc = conv_wgrad(c, T, L1, B, gY, L1.x, s))) return rc;
        const bool first_block_of_net = i1 == 1;
        if (!first_block_of_net || true) {
            if ((rc = conv_dgrad(c, T, L1, B, gY, gA, gUp, gLow, false, s))) return rc;      // gA <- d X (main branch)
        }
        // identity branch: dS (gB) flows to X directly, or through the downsample convolution + its BatchNorm
        const long long n_x = (long long)B * L1.in_hw * L1.in_hw * L1.cin;
        if (has_ds) {
            TrainLayer& Ld = T->L[id];
            if ((rc = bn_backward(c, T, Ld, M3, gB, false, gY, nullptr, s))) return rc;
            if ((rc = conv_wgrad(c, T, Ld, B, gY, Ld.x, s))) return rc;
            if ((rc = conv_dgrad(c, T, Ld, B, gY, gA, gUp, gLow, true, s))) return rc;       // gA += d X (downsample branch)
        } else {
            add_inplace_kernel<<<n4(n_x), 256, 0, s>>>(gA, gB, n_x / 4);
            c->launches += 1;
        }
        i = i1 - 1;
    }
    {   // stem: max-pool, ReLU + bn1, the two 7x7 convolutions' weights (no data gradient: the inputs need none)
        TrainLayer& L = T->L[0];
        const long long tot = (long long)B * 128 * 128 * 64;
        maxpool_bwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(gA, T->pool_idx, gB, tot);
        c->launches += 1;
        if ((rc = bn_backward(c, T, L, (long long)B * 128 * 128, gB, true, gY, nullptr, s))) return rc;
        if ((rc = conv_wgrad(c, T, L, B, gY, L.x, s))) return rc;
    }
    if (apply && (rc = train_apply(c, lr, momentum, wd, s))) return rc;
    IVOSW_CUDA(cudaStreamSynchronize(s));
    T->steps += 1;
    return IVOSW_OK;
}

// clamp + SGD (:266-269) on the current step's gradient (possibly all-reduced across ranks first), per trainable range
int train_apply(ivosw_ctx* c, float lr, float momentum, float wd, cudaStream_t s) {
    TrainState* T = static_cast<TrainState*>(c->train_state);
    if (!T) { set_error("ivosw_assess_train_begin has not been called"); return IVOSW_ERR_STATE; }
    for (size_t li = 0; li < T->L.size(); ++li) {
        const TrainLayer& L = T->L[li];
        const long long n = (long long)(L.rm_off - L.w_off);          // weight, gamma, beta are contiguous
        sgd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(T->blob + L.w_off, T->gnew + L.w_off, T->gacc + L.w_off,
                                                               T->mom + L.w_off, n, lr, momentum, wd, 1);
    }
    sgd_kernel<<<(2049 + 255) / 256, 256, 0, s>>>(T->blob + T->fc_off, T->gnew + T->fc_off, T->gacc + T->fc_off, T->mom + T->fc_off,
                                                  2049, lr, momentum, wd, 1);
    c->launches += (long long)T->L.size() + 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int train_export(ivosw_ctx* c, float* blob_host, float* grad_host, cudaStream_t s) {
    TrainState* T = static_cast<TrainState*>(c->train_state);
    if (!T) { set_error("ivosw_assess_train_begin has not been called"); return IVOSW_ERR_STATE; }
    if (blob_host) IVOSW_CUDA(cudaMemcpyAsync(blob_host, T->blob, sizeof(float) * T->n_blob, cudaMemcpyDeviceToHost, s));
    if (grad_host) IVOSW_CUDA(cudaMemcpyAsync(grad_host, T->gacc, sizeof(float) * T->n_blob, cudaMemcpyDeviceToHost, s));
    IVOSW_CUDA(cudaStreamSynchronize(s));
    return IVOSW_OK;
}

int train_grad_buffer(ivosw_ctx* c, float** gnew_dev, size_t* n) {
    TrainState* T = static_cast<TrainState*>(c->train_state);
    if (!T) { set_error("ivosw_assess_train_begin has not been called"); return IVOSW_ERR_STATE; }
    *gnew_dev = T->gnew; *n = T->n_blob;
    return IVOSW_OK;
}

}  // namespace ivosw
