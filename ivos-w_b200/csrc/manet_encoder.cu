// The MANet feature extractor (IntVOS.extract_feature, call site eval_agent_manet.py:316-328) on B200:
// DeepLabv3+ ResNet-101 at output stride 16, ASPP (256), shortcut decoder (48), semantic-embedding head (100 channels
// at 1/4 resolution) — SURVEY.md §2.1 K8.
//
// RESTATEMENT, PARITY UNPINNED: the network's source is not part of the reference tree (README.md:38-40 clones an
// unpinned third-party repository); the architecture follows ivosw/manet_arch.py (published DeepLabv3+ / FEELVOS / MANet
// descriptions under the hyper-parameters of utils/config_manet/config.py:108-120) and is checked against
// oracle/manet_encoder_ref.py, a restatement of the same description.
//
// All 111 dense convolutions run on the tcgen05 implicit-GEMM kernel of conv_tc.cu (split-fp16, fp32-grade) through
// launch_conv_tc_g: feature maps of any size live on power-of-two CANVASES (480 x 854: 120 x 214 on 128 x 256, 60 x 107 on
// 64 x 128, 30 x 54 on 32 x 64) whose positions outside the map are kept at zero — they ARE the zero padding of the next
// (possibly dilated) convolution, TMA's out-of-bounds fill supplies the rest — so a 128-pixel GEMM tile is still whole canvas
// rows and one TMA box per filter tap.  The downsample branch of each stage's first bottleneck is fused into its conv3 as
// one concatenated-K GEMM; the four ASPP branches and the decoder's two inputs are written straight into channel slices
// of their concatenation buffers (TMA store maps with a row pitch), so no concatenation pass exists.
// CUDA-core kernels only where there is no GEMM: the 3-channel 7x7 stem + max-pool, ASPP's image-pooling branch (a
// 2048-long mean and a 2048 x 256 GEMV per frame), the bilinear x4 upsampling, the depthwise 3x3 of the embedding head.
// Canvas efficiency: 79 % of the issued MMA work lands inside the feature maps at 480 x 854.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ivosw_internal.h"

namespace ivosw {

namespace {

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float ld_split(const __half* hi, const __half* lo, size_t i, bool use_lo) {
    const float h = __half2float(hi[i]);
    return use_lo ? fmaf(__half2float(lo[i]), 1.0f / 2048.0f, h) : h;
}
__device__ __forceinline__ void st_split(__half* hi, __half* lo, size_t i, float v) {
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn((v - __half2float(h)) * 2048.0f);
}

static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ------------------------------------------------------------------------------------------------ stem
constexpr int ES_TH = 8, ES_TW = 16;                            // output tile
constexpr int ES_PH = ES_TH * 2 + 5, ES_PW = ES_TW * 2 + 5;     // 21 x 37 input patch
constexpr int ES_K = 147;                                       // 7 * 7 * 3

// 7x7 stride-2 pad-3 convolution 3 -> 64 + BatchNorm + ReLU: frames [B][3][H][W] fp32 -> c1 [B][h2][w2][64] fp32
__global__ void __launch_bounds__(256) enc_stem_conv_kernel(const float* __restrict__ frames, const float* __restrict__ wgt /*[147][64]*/,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            float* __restrict__ c1, int H, int W, int h2, int w2) {
    extern __shared__ float smem[];
    float* sw = smem;                                   // 147 * 64
    float* sp = smem + ES_K * 64;                       // 3 * 21 * 37
    const int b = blockIdx.z;
    const int oy0 = blockIdx.y * ES_TH, ox0 = blockIdx.x * ES_TW;
    for (int i = threadIdx.x; i < ES_K * 64; i += 256) sw[i] = __ldg(wgt + i);
    const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
    for (int i = threadIdx.x; i < 3 * ES_PH * ES_PW; i += 256) {
        const int ch = i / (ES_PH * ES_PW), r = i - ch * (ES_PH * ES_PW);
        const int py = r / ES_PW, px = r - py * ES_PW;
        const int iy = iy0 + py, ix = ix0 + px;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(frames + (((size_t)b * 3 + ch) * H + iy) * W + ix);
        sp[i] = v;
    }
    __syncthreads();
    const int cg = threadIdx.x & 15;                    // channels 4*cg .. 4*cg+3
    const int px = threadIdx.x >> 4;                    // tile column
    float acc[ES_TH][4];
#pragma unroll
    for (int r = 0; r < ES_TH; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    for (int kh = 0; kh < 7; ++kh)
        for (int kw = 0; kw < 7; ++kw)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float4 w4 = *reinterpret_cast<const float4*>(sw + ((kh * 7 + kw) * 3 + ci) * 64 + cg * 4);
#pragma unroll
                for (int r = 0; r < ES_TH; ++r) {
                    const float v = sp[ci * (ES_PH * ES_PW) + (r * 2 + kh) * ES_PW + px * 2 + kw];
                    acc[r][0] = fmaf(v, w4.x, acc[r][0]); acc[r][1] = fmaf(v, w4.y, acc[r][1]);
                    acc[r][2] = fmaf(v, w4.z, acc[r][2]); acc[r][3] = fmaf(v, w4.w, acc[r][3]);
                }
            }
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + cg), sh = __ldg(reinterpret_cast<const float4*>(shift) + cg);
    const int ox = ox0 + px;
    if (ox >= w2) return;
#pragma unroll
    for (int r = 0; r < ES_TH; ++r) {
        const int oy = oy0 + r;
        if (oy >= h2) continue;
        float4 o;
        o.x = fmaxf(fmaf(acc[r][0], sc.x, sh.x), 0.f); o.y = fmaxf(fmaf(acc[r][1], sc.y, sh.y), 0.f);
        o.z = fmaxf(fmaf(acc[r][2], sc.z, sh.z), 0.f); o.w = fmaxf(fmaf(acc[r][3], sc.w, sh.w), 0.f);
        reinterpret_cast<float4*>(c1 + (((size_t)b * h2 + oy) * w2 + ox) * 64)[cg] = o;
    }
}

// 3x3 stride-2 pad-1 max-pool of c1 -> split planes on the 1/4 canvas (zeros outside the feature map)
__global__ void __launch_bounds__(256) enc_pool_kernel(const float* __restrict__ c1, __half* __restrict__ hi, __half* __restrict__ lo,
                                                       int h2, int w2, int h4, int w4, int Hp, int Wp, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B * Hp * Wp * 64
    if (i >= total) return;
    const int c = (int)(i & 63);
    long long p = i >> 6;
    const int x = (int)(p % Wp); p /= Wp;
    const int y = (int)(p % Hp);
    const long long b = p / Hp;
    float m = 0.f;
    if (y < h4 && x < w4) {
        m = -INFINITY;
        for (int dy = -1; dy <= 1; ++dy) {
            const int iy = y * 2 + dy;
            if (iy < 0 || iy >= h2) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int ix = x * 2 + dx;
                if (ix < 0 || ix >= w2) continue;
                m = fmaxf(m, c1[((b * h2 + iy) * w2 + ix) * 64 + c]);
            }
        }
    }
    st_split(hi, lo, (size_t)i, m);
}

// 8 consecutive channels of a split-fp16 tensor <-> 8 floats (one 16-byte load / store per plane)
__device__ __forceinline__ void ld_split8(const __half* hi, const __half* lo, size_t i, bool use_lo, float* v) {
    const uint4 h = *reinterpret_cast<const uint4*>(hi + i);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
    uint32_t lw[4] = {0u, 0u, 0u, 0u};
    if (use_lo) { const uint4 l = *reinterpret_cast<const uint4*>(lo + i); lw[0] = l.x; lw[1] = l.y; lw[2] = l.z; lw[3] = l.w; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[u]));
        const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[u]));
        v[2 * u] = fmaf(lf.x, 1.0f / 2048.0f, hf.x); v[2 * u + 1] = fmaf(lf.y, 1.0f / 2048.0f, hf.y);
    }
}
__device__ __forceinline__ void st_split8(__half* hi, __half* lo, size_t i, const float* v) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float a = fminf(fmaxf(v[2 * u], -65504.f), 65504.f), b = fminf(fmaxf(v[2 * u + 1], -65504.f), 65504.f);
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
        hw[u] = *reinterpret_cast<const uint32_t*>(&h); lw[u] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(hi + i) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo + i) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// ------------------------------------------------------------------------------------------------ ASPP image pooling
// partial sums over a slice of the canvas (zeros outside the map, so canvas sums are map sums):
// x [B][px][C] split planes -> part[b][split][C] fp32; 8 channels per thread
constexpr int GAP_SPLITS = 16;
__global__ void __launch_bounds__(256) enc_gap_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int use_lo, int px,
                                                      int C, float* __restrict__ part) {
    const int c8 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int b = blockIdx.y, sp = blockIdx.z;
    if (c8 >= C) return;
    const int per = (px + GAP_SPLITS - 1) / GAP_SPLITS;
    const int p0 = sp * per, p1 = min(px, p0 + per);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = p0; p < p1; ++p) {
        float v[8];
        ld_split8(hi, lo, ((size_t)b * px + p) * C + c8, use_lo, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[((size_t)b * GAP_SPLITS + sp) * C + c8 + k] = acc[k];
}

// pooled vector (fixed-order sum of the partials) -> 1x1 convolution + BatchNorm + ReLU: val[b][256]
__global__ void __launch_bounds__(256) enc_gap_gemv_kernel(const float* __restrict__ part, float inv_n, const float* __restrict__ w /*[256][2048]*/,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           float* __restrict__ val) {
    __shared__ float g[2048];
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < 2048; c += 256) {
        float s = 0.f;
        for (int k = 0; k < GAP_SPLITS; ++k) s += part[((size_t)b * GAP_SPLITS + k) * 2048 + c];
        g[c] = s * inv_n;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int co = warp; co < 256; co += 8) {                      // one warp per output channel: coalesced weight rows
        const float* wr = w + (size_t)co * 2048;
        float a = 0.f;
        for (int k = lane; k < 2048; k += 32) a = fmaf(wr[k], g[k], a);
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) val[b * 256 + co] = fmaxf(fmaf(a, scale[co], shift[co]), 0.f);
    }
}

// broadcast over the feature map (the bilinear upsampling of a 1 x 1 map is a constant) into channels
// [coff, coff + 256) of the concatenation buffer (row pitch ld); 8 channels per thread
__global__ void __launch_bounds__(256) enc_gap_broadcast_kernel(const float* __restrict__ val, __half* __restrict__ hi, __half* __restrict__ lo,
                                                                int h, int w_, int Hp, int Wp, int ld, int coff, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B * Hp * Wp * 32
    if (i >= total) return;
    const int c8 = (int)(i & 31) * 8;
    long long p = i >> 5;
    const int x = (int)(p % Wp); p /= Wp;
    const int y = (int)(p % Hp);
    const long long b = p / Hp;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (y < h && x < w_) ? val[b * 256 + c8 + k] : 0.f;
    st_split8(hi, lo, ((size_t)(b * Hp + y) * Wp + x) * ld + coff + c8, v);
}

// ------------------------------------------------------------------------------------------------ decoder pieces
// bilinear upsampling (align_corners=True) of a C-channel map from canvas (Hs, Ws; map hs x ws) to canvas (Hd, Wd; map
// hd x wd), into channels [coff, coff + C) of a buffer with row pitch ld; 8 channels per thread
__global__ void __launch_bounds__(256) enc_upsample_kernel(const __half* __restrict__ shi, const __half* __restrict__ slo, int use_lo,
                                                           int hs, int ws, int Hs, int Ws, int hd, int wd, int Hd, int Wd, int C,
                                                           float sy, float sx, __half* __restrict__ dhi, __half* __restrict__ dlo,
                                                           int ld, int coff, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B * Hd * Wd * C / 8
    if (i >= total) return;
    const int c8n = C / 8;
    const int c8 = (int)(i % c8n) * 8;
    long long p = i / c8n;
    const int x = (int)(p % Wd); p /= Wd;
    const int y = (int)(p % Hd);
    const long long b = p / Hd;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (y < hd && x < wd) {
        // ATen area_pixel_compute_source_index(align_corners=True): src = scale * dst
        const float fy = sy * (float)y, fx = sx * (float)x;
        const int y0 = min((int)fy, hs - 1), x0 = min((int)fx, ws - 1);
        const int y1 = y0 + (y0 < hs - 1 ? 1 : 0), x1 = x0 + (x0 < ws - 1 ? 1 : 0);
        const float ly1 = fy - (float)y0, ly0 = 1.f - ly1, lx1 = fx - (float)x0, lx0 = 1.f - lx1;
        const size_t base = (size_t)b * Hs * Ws;
        float v00[8], v01[8], v10[8], v11[8];
        ld_split8(shi, slo, (base + (size_t)y0 * Ws + x0) * C + c8, use_lo, v00);
        ld_split8(shi, slo, (base + (size_t)y0 * Ws + x1) * C + c8, use_lo, v01);
        ld_split8(shi, slo, (base + (size_t)y1 * Ws + x0) * C + c8, use_lo, v10);
        ld_split8(shi, slo, (base + (size_t)y1 * Ws + x1) * C + c8, use_lo, v11);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ly0 * (lx0 * v00[k] + lx1 * v01[k]) + ly1 * (lx0 * v10[k] + lx1 * v11[k]);
    }
    st_split8(dhi, dlo, ((size_t)(b * Hd + y) * Wd + x) * ld + coff + c8, v);
}

// depthwise 3x3 (pad 1) + BatchNorm + ReLU on a canvas (zeros outside the map = the padding); w [C][3][3];
// 8 channels per thread
__global__ void __launch_bounds__(256) enc_dwconv_kernel(const __half* __restrict__ shi, const __half* __restrict__ slo, int use_lo,
                                                         const float* __restrict__ w, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, int h, int w_, int Hp, int Wp, int C,
                                                         __half* __restrict__ dhi, __half* __restrict__ dlo, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B * Hp * Wp * C / 8
    if (i >= total) return;
    const int c8n = C / 8;
    const int c8 = (int)(i % c8n) * 8;
    long long p = i / c8n;
    const int x = (int)(p % Wp); p /= Wp;
    const int y = (int)(p % Hp);
    const long long b = p / Hp;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (y < h && x < w_) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int dy = -1; dy <= 1; ++dy) {
            const int iy = y + dy;
            if (iy < 0 || iy >= Hp) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int ix = x + dx;
                if (ix < 0 || ix >= Wp) continue;
                float t[8];
                ld_split8(shi, slo, ((size_t)(b * Hp + iy) * Wp + ix) * C + c8, use_lo, t);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(t[k], __ldg(w + (c8 + k) * 9 + (dy + 1) * 3 + dx + 1), acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(acc[k], scale[c8 + k], shift[c8 + k]), 0.f);
    }
    st_split8(dhi, dlo, ((size_t)(b * Hp + y) * Wp + x) * C + c8, v);
}

// embedding [B][Hp*Wp][Cp] split planes -> [B][C][h][w] fp32 (the layout extract_feature returns)
__global__ void __launch_bounds__(256) enc_export_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int use_lo, int h,
                                                         int w_, int Hp, int Wp, int Cp, int C, float* __restrict__ out, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B * C * h * w
    if (i >= total) return;
    const int x = (int)(i % w_);
    long long p = i / w_;
    const int y = (int)(p % h); p /= h;
    const int c = (int)(p % C);
    const long long b = p / C;
    out[i] = ld_split(hi, lo, ((size_t)(b * Hp + y) * Wp + x) * Cp + c, use_lo);
}

// ------------------------------------------------------------------------------------------------ graph
struct EncConvW {               // host-side description of one blob entry (mirrors ivosw/manet_arch.py::convs())
    int cin, cout, k, stride, dil, groups, relu;
    size_t w_off, bn_off;       // offsets into the blob (floats): weight [cout][k][k][cin/groups], then gamma, beta, mean, var
};

struct EncState {
    bool loaded = false;
    std::vector<EncConvW> W;                    // 115 entries, blob order
    std::vector<float> blob;                    // host copy (weights are re-packed per geometry-independent layer once)
    // device parameters
    float* stem_w = nullptr;                    // [147][64]
    float *stem_scale = nullptr, *stem_shift = nullptr;
    std::vector<GConv> convs;                   // tensor-core convolutions in execution order (geometry filled per call)
    std::vector<int> conv_src;                  // blob index of each (fused tails: the conv3 entry)
    float *gap_w = nullptr, *gap_scale = nullptr, *gap_shift = nullptr;
    float *dw_w = nullptr, *dw_scale = nullptr, *dw_shift = nullptr;
    std::vector<void*> owned;                   // device allocations to free
    DeviceBuffer arena;
};

static void fold(const float* bn, int n, std::vector<float>& sc, std::vector<float>& sh) {
    sc.resize(n); sh.resize(n);
    for (int i = 0; i < n; ++i) {
        const double s = (double)bn[i] / std::sqrt((double)bn[3 * n + i] + (double)BN_EPS);
        sc[i] = (float)s;
        sh[i] = (float)((double)bn[n + i] - (double)bn[2 * n + i] * s);
    }
}

template <class T>
static int dev_upload(EncState* E, T** dst, const std::vector<T>& src) {
    IVOSW_CUDA(cudaMalloc(dst, src.size() * sizeof(T)));
    E->owned.push_back(*dst);
    IVOSW_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return IVOSW_OK;
}

// weight of one tensor-core convolution: [cout_p][k*k*cin_p (+ cin2)] fp32 -> split planes; `sc` (optional) folds a
// BatchNorm scale into the rows; channels beyond the real cin / cout are zero
static int upload_conv(EncState* E, GConv& G, const float* w, int cout, int cin, int k, const float* sc, const float* w2, int cin2,
                       const float* sc2) {
    const int cout_p = G.cout, cin_p = G.cin, K = k * k * cin_p + cin2;
    std::vector<__half> hi((size_t)cout_p * K, __float2half_rn(0.f)), lo((size_t)cout_p * K, __float2half_rn(0.f));
    for (int o = 0; o < cout; ++o) {
        for (int t = 0; t < k * k; ++t)
            for (int ci = 0; ci < cin; ++ci) {
                const double v = (double)w[((size_t)o * k * k + t) * cin + ci] * (sc ? (double)sc[o] : 1.0);
                const float f = (float)v;
                const __half h = __float2half_rn(f);
                hi[(size_t)o * K + (size_t)t * cin_p + ci] = h;
                lo[(size_t)o * K + (size_t)t * cin_p + ci] = __float2half_rn((f - __half2float(h)) * 2048.0f);
            }
        for (int ci = 0; ci < cin2; ++ci) {
            const float f = (float)((double)w2[(size_t)o * cin2 + ci] * (double)sc2[o]);
            const __half h = __float2half_rn(f);
            hi[(size_t)o * K + (size_t)k * k * cin_p + ci] = h;
            lo[(size_t)o * K + (size_t)k * k * cin_p + ci] = __float2half_rn((f - __half2float(h)) * 2048.0f);
        }
    }
    int rc;
    if ((rc = dev_upload(E, &G.w_hi, hi))) return rc;
    return dev_upload(E, &G.w_lo, lo);
}

static int pad64(int c) { return (c + 63) / 64 * 64; }

}  // namespace

void manet_encoder_release(ivosw_ctx* c) {
    EncState* E = static_cast<EncState*>(c->enc_state);
    if (!E) return;
    for (void* p : E->owned) cudaFree(p);
    release(E->arena);
    delete E;
    c->enc_state = nullptr;
}

size_t manet_encoder_blob_floats() {
    // mirrors ivosw/manet_arch.py::convs(): weight + 4 BatchNorm vectors per entry
    size_t n = (size_t)64 * 49 * 3 + 4 * 64;
    const int planes_[4] = {64, 128, 256, 512}, blocks_[4] = {3, 4, 23, 3};
    int inpl = 64;
    for (int s = 0; s < 4; ++s)
        for (int b = 0; b < blocks_[s]; ++b) {
            const int p = planes_[s];
            n += (size_t)p * inpl + 4 * p + (size_t)p * 9 * p + 4 * p + (size_t)4 * p * p + 16 * p;
            if (b == 0) n += (size_t)4 * p * inpl + 16 * p;
            inpl = 4 * p;
        }
    n += (size_t)256 * 2048 + 1024 + 3 * ((size_t)256 * 9 * 2048 + 1024) + (size_t)256 * 2048 + 1024 + (size_t)256 * 1280 + 1024;
    n += (size_t)48 * 256 + 192 + (size_t)256 * 9 * 304 + 1024 + (size_t)256 * 9 * 256 + 1024 + (size_t)256 * 9 + 1024 +
         (size_t)100 * 256 + 400;
    return n;
}

int manet_encoder_load(ivosw_ctx* c, const float* blob, size_t n_floats) {
    if (n_floats != manet_encoder_blob_floats()) { set_error("invalid argument: MANet encoder blob length"); return IVOSW_ERR_INVALID; }
    manet_encoder_release(c);
    EncState* E = new EncState();
    c->enc_state = E;
    int rc;
    size_t off = 0;
    auto entry = [&](int cin, int cout, int k, int stride, int dil, int groups, int relu) {
        EncConvW w{cin, cout, k, stride, dil, groups, relu, off, 0};
        off += (size_t)cout * k * k * (cin / groups);
        w.bn_off = off;
        off += 4 * (size_t)cout;
        E->W.push_back(w);
        return (int)E->W.size() - 1;
    };
    std::vector<float> sc, sh, sc2, sh2;
    {   // stem: [64][7][7][3] -> [(kh, kw, ci)][64]
        const int i0 = entry(3, 64, 7, 2, 1, 1, 1);
        const EncConvW& w = E->W[i0];
        std::vector<float> wt((size_t)ES_K * 64);
        for (int o = 0; o < 64; ++o)
            for (int k = 0; k < ES_K; ++k) wt[(size_t)k * 64 + o] = blob[w.w_off + (size_t)o * ES_K + k];
        fold(blob + w.bn_off, 64, sc, sh);
        if ((rc = dev_upload(E, &E->stem_w, wt)) || (rc = dev_upload(E, &E->stem_scale, sc)) || (rc = dev_upload(E, &E->stem_shift, sh))) return rc;
    }
    auto add_conv = [&](int wi, int wi_ds) -> int {
        // one tensor-core convolution from blob entry wi (conv + bn), optionally fused with the downsample entry wi_ds
        const EncConvW& w = E->W[wi];
        GConv G;
        G.cin = pad64(w.cin); G.cout = pad64(w.cout); G.k = w.k; G.stride = w.stride; G.dil = w.dil; G.relu = w.relu;
        fold(blob + w.bn_off, w.cout, sc, sh);
        std::vector<float> scale(G.cout, 1.0f), shift(G.cout, 0.0f);
        int r;
        if (wi_ds >= 0) {
            const EncConvW& d = E->W[wi_ds];
            fold(blob + d.bn_off, d.cout, sc2, sh2);
            for (int o = 0; o < w.cout; ++o) shift[o] = (float)((double)sh[o] + (double)sh2[o]);
            G.cin2 = d.cin; G.stride2 = d.stride;
            if ((r = upload_conv(E, G, blob + w.w_off, w.cout, w.cin, w.k, sc.data(), blob + d.w_off, d.cin, sc2.data()))) return r;
        } else {
            for (int o = 0; o < w.cout; ++o) { scale[o] = sc[o]; shift[o] = sh[o]; }
            if ((r = upload_conv(E, G, blob + w.w_off, w.cout, w.cin, w.k, nullptr, nullptr, 0, nullptr))) return r;
        }
        if ((r = dev_upload(E, &G.scale, scale)) || (r = dev_upload(E, &G.shift, shift))) return r;
        E->convs.push_back(G);
        E->conv_src.push_back(wi);
        return IVOSW_OK;
    };
    const int planes_[4] = {64, 128, 256, 512}, blocks_[4] = {3, 4, 23, 3}, stride_[4] = {1, 2, 2, 1};
    const int dil4[3] = {2, 4, 8};
    int inpl = 64;
    for (int s = 0; s < 4; ++s)
        for (int b = 0; b < blocks_[s]; ++b) {
            const int p = planes_[s], st = b == 0 ? stride_[s] : 1, d = s == 3 ? dil4[b] : 1;
            const int i1 = entry(inpl, p, 1, 1, 1, 1, 1);
            const int i2 = entry(p, p, 3, st, d, 1, 1);
            const int id = b == 0 ? entry(inpl, 4 * p, 1, st, 1, 1, 0) : -1;
            const int i3 = entry(p, 4 * p, 1, 1, 1, 1, 1);
            if ((rc = add_conv(i1, -1)) || (rc = add_conv(i2, -1)) || (rc = add_conv(i3, id))) return rc;
            inpl = 4 * p;
        }
    {   // ASPP: 1x1, three dilated 3x3, image pooling (CUDA cores), 1x1 over the concatenation
        const int a1 = entry(2048, 256, 1, 1, 1, 1, 1);
        const int a2 = entry(2048, 256, 3, 1, 6, 1, 1), a3 = entry(2048, 256, 3, 1, 12, 1, 1), a4 = entry(2048, 256, 3, 1, 18, 1, 1);
        const int ag = entry(2048, 256, 1, 1, 1, 1, 1);
        const int ac = entry(1280, 256, 1, 1, 1, 1, 1);
        if ((rc = add_conv(a1, -1)) || (rc = add_conv(a2, -1)) || (rc = add_conv(a3, -1)) || (rc = add_conv(a4, -1))) return rc;
        const EncConvW& g = E->W[ag];
        std::vector<float> gw(blob + g.w_off, blob + g.w_off + (size_t)256 * 2048);
        fold(blob + g.bn_off, 256, sc, sh);
        if ((rc = dev_upload(E, &E->gap_w, gw)) || (rc = dev_upload(E, &E->gap_scale, sc)) || (rc = dev_upload(E, &E->gap_shift, sh))) return rc;
        if ((rc = add_conv(ac, -1))) return rc;
    }
    {   // decoder + embedding head
        const int d1 = entry(256, 48, 1, 1, 1, 1, 1), l0 = entry(304, 256, 3, 1, 1, 1, 1), l4 = entry(256, 256, 3, 1, 1, 1, 1);
        const int dw = entry(256, 256, 3, 1, 1, 256, 1), pw = entry(256, 100, 1, 1, 1, 1, 1);
        if ((rc = add_conv(d1, -1)) || (rc = add_conv(l0, -1)) || (rc = add_conv(l4, -1))) return rc;
        const EncConvW& w = E->W[dw];
        std::vector<float> ww(blob + w.w_off, blob + w.w_off + (size_t)256 * 9);
        fold(blob + w.bn_off, 256, sc, sh);
        if ((rc = dev_upload(E, &E->dw_w, ww)) || (rc = dev_upload(E, &E->dw_scale, sc)) || (rc = dev_upload(E, &E->dw_shift, sh))) return rc;
        if ((rc = add_conv(pw, -1))) return rc;
    }
    if (off != n_floats) { set_error("MANet encoder blob walk does not match its length"); return IVOSW_ERR_STATE; }
    IVOSW_CUDA(cudaDeviceSynchronize());
    E->loaded = true;
    return IVOSW_OK;
}

// frames [B][3][H][W] fp32 (normalised, device) -> embedding [B][100][h4][w4] fp32 (device)
int manet_encoder_forward(ivosw_ctx* c, const float* frames, int B, int H, int W, float* out, int terms, cudaStream_t s) {
    EncState* E = static_cast<EncState*>(c->enc_state);
    if (!E || !E->loaded) { set_error("MANet encoder weights not loaded"); return IVOSW_ERR_STATE; }
    auto half_ = [](int n) { return (n + 2 - 3) / 2 + 1; };
    const int h2 = (H + 6 - 7) / 2 + 1, w2 = (W + 6 - 7) / 2 + 1;
    const int h4 = half_(h2), w4 = half_(w2), h8 = half_(h4), w8 = half_(w4), h16 = half_(h8), w16 = half_(w8);
    // canvases: powers of two, at least 8 x 16 (one 128-pixel GEMM tile)
    const int H4 = std::max(8, pow2ceil(h4)), W4 = std::max(16, pow2ceil(w4));
    const int H8 = std::max(8, pow2ceil(h8)), W8 = std::max(16, pow2ceil(w8));
    const int H16 = std::max(8, pow2ceil(h16)), W16 = std::max(16, pow2ceil(w16));
    if (2 * H8 != H4 || 2 * W8 != W4 || 2 * H16 != H8 || 2 * W16 != W8) {
        // (a stride-2 layer reads its input canvas through a parity view: the canvases must nest exactly)
        set_error("invalid argument: frame size too small for the encoder's canvases (needs >= 113 x 241 or so)");
        return IVOSW_ERR_INVALID;
    }
    const int use_lo = terms == 3;
    const size_t P4 = (size_t)H4 * W4, P8 = (size_t)H8 * W8, P16 = (size_t)H16 * W16;
    // ---- workspace (split planes: bytes = pixels * channels * 4)
    size_t total = 0;
    auto take = [&](size_t bytes) { size_t o = total; total += (bytes + 1023) & ~(size_t)1023; return o; };
    auto planes = [&](size_t px, int C) { return take((size_t)B * px * C * 4); };
    const size_t o_c1 = take((size_t)B * h2 * w2 * 64 * 4);
    const size_t o_p4 = planes(P4, 64), o_t1_4 = planes(P4, 128), o_t2_4 = planes(P4, 64), o_x4a = planes(P4, 256), o_x4b = planes(P4, 256);
    const size_t o_t1_8 = planes(P8, 256), o_t2_8 = planes(P8, 128), o_x8a = planes(P8, 512), o_x8b = planes(P8, 512);
    const size_t o_t1_16 = planes(P16, 512), o_t2_16 = planes(P16, 512), o_x16a = planes(P16, 2048), o_x16b = planes(P16, 2048);
    const size_t o_cat16 = planes(P16, 1280), o_a16 = planes(P16, 256), o_gap = take((size_t)B * (GAP_SPLITS * 2048 + 256) * 4);
    const size_t o_cat4 = planes(P4, 320), o_d4a = planes(P4, 256), o_d4b = planes(P4, 256), o_e4 = planes(P4, 128);
    total += (size_t)64 << 20;      // slack: an 8 x 16 canvas tile of a missing second image, as in conv_tc.cu
    int rc;
    if ((rc = ensure(E->arena, total))) return rc;
    char* base = (char*)E->arena.p;
    auto view = [&](size_t off, size_t px, int C) {
        return SplitAct{reinterpret_cast<__half*>(base + off), reinterpret_cast<__half*>(base + off + (size_t)B * px * C * 2)};
    };
    // ---- stem + max-pool
    {
        const size_t smem = ((size_t)ES_K * 64 + 3 * ES_PH * ES_PW) * sizeof(float);
        static bool attr[64] = {};
        if (!attr[c->device & 63]) {
            IVOSW_CUDA(cudaFuncSetAttribute(enc_stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr[c->device & 63] = true;
        }
        dim3 grid((w2 + ES_TW - 1) / ES_TW, (h2 + ES_TH - 1) / ES_TH, B);
        enc_stem_conv_kernel<<<grid, 256, smem, s>>>(frames, E->stem_w, E->stem_scale, E->stem_shift, (float*)(base + o_c1), H, W, h2, w2);
        const SplitAct p4 = view(o_p4, P4, 64);
        const long long tot = (long long)B * P4 * 64;
        enc_pool_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>((const float*)(base + o_c1), p4.hi, p4.lo, h2, w2, h4, w4, H4, W4, tot);
        c->launches += 2;
        IVOSW_CUDA(cudaGetLastError());
    }
    // ---- backbone: 33 bottlenecks = 99 tensor-core convolutions (downsample branches fused into conv3)
    struct Res { int h, w, Hp, Wp; size_t px; };
    const Res R4{h4, w4, H4, W4, P4}, R8{h8, w8, H8, W8, P8}, R16{h16, w16, H16, W16, P16};
    const int planes_[4] = {64, 128, 256, 512}, blocks_[4] = {3, 4, 23, 3}, stride_[4] = {1, 2, 2, 1};
    const Res* res_in[4] = {&R4, &R4, &R8, &R16};
    const Res* res_out[4] = {&R4, &R8, &R16, &R16};
    // per stage: (t1 buffer for block 0 at the INPUT resolution, t1/t2 at the output resolution, two block outputs)
    const size_t t1_first[4] = {o_t1_4, o_t1_4, o_t1_8, o_t1_16}, t1_rest[4] = {o_t1_4, o_t1_8, o_t1_16, o_t1_16};
    const size_t t2_[4] = {o_t2_4, o_t2_8, o_t2_16, o_t2_16};
    const size_t xa_[4] = {o_x4a, o_x8a, o_x16a, o_x16a}, xb_[4] = {o_x4b, o_x8b, o_x16b, o_x16b};
    SplitAct x = view(o_p4, P4, 64);
    int x_c = 64;
    size_t ci = 0;
    SplitAct low4{nullptr, nullptr};
    for (int st = 0; st < 4; ++st) {
        const int p = planes_[st];
        // layer3 writes 1024 channels into the 2048-channel buffers' space and layer4 2048: alternate a / b, and make sure
        // the stage's last block does not overwrite the buffer layer1's output (the decoder's low-level input) lives in
        int flip = 0;
        if (st == 3) flip = ((blocks_[2] & 1) ? 1 : 0);              // layer4 continues the a / b alternation of layer3
        for (int b = 0; b < blocks_[st]; ++b) {
            const Res& Ri = b == 0 ? *res_in[st] : *res_out[st];
            const Res& Ro = *res_out[st];
            GConv g1 = E->convs[ci], g2 = E->convs[ci + 1], g3 = E->convs[ci + 2];
            ci += 3;
            const SplitAct t1 = view(b == 0 ? t1_first[st] : t1_rest[st], Ri.px, p);
            const SplitAct t2 = view(t2_[st], Ro.px, p);
            const SplitAct y = view(flip ? xb_[st] : xa_[st], Ro.px, 4 * p);
            g1.in_hp = Ri.Hp; g1.in_wp = Ri.Wp; g1.out_hp = Ri.Hp; g1.out_wp = Ri.Wp; g1.valid_h = Ri.h; g1.valid_w = Ri.w;
            if ((rc = launch_conv_tc_g(c, g1, x, nullptr, nullptr, t1, g1.cout, B, terms, s))) return rc;
            g2.in_hp = Ri.Hp; g2.in_wp = Ri.Wp; g2.out_hp = Ro.Hp; g2.out_wp = Ro.Wp; g2.valid_h = Ro.h; g2.valid_w = Ro.w;
            if ((rc = launch_conv_tc_g(c, g2, t1, nullptr, nullptr, t2, g2.cout, B, terms, s))) return rc;
            g3.in_hp = Ro.Hp; g3.in_wp = Ro.Wp; g3.out_hp = Ro.Hp; g3.out_wp = Ro.Wp; g3.valid_h = Ro.h; g3.valid_w = Ro.w;
            if (g3.cin2) {
                g3.in2_hp = Ri.Hp; g3.in2_wp = Ri.Wp;
                if ((rc = launch_conv_tc_g(c, g3, t2, &x, nullptr, y, g3.cout, B, terms, s))) return rc;
            } else if ((rc = launch_conv_tc_g(c, g3, t2, nullptr, &x, y, g3.cout, B, terms, s))) return rc;
            x = y; x_c = 4 * p;
            flip ^= 1;
        }
        if (st == 0) {
            low4 = x;
            // layer1 ends in x4a (3 blocks: a, b, a); layer2 only writes the 1/8 buffers, so it stays intact
        }
    }
    (void)x_c;
    // ---- ASPP
    const SplitAct cat16 = view(o_cat16, P16, 1280);
    for (int j = 0; j < 4; ++j) {
        GConv g = E->convs[ci++];
        g.in_hp = H16; g.in_wp = W16; g.out_hp = H16; g.out_wp = W16; g.valid_h = h16; g.valid_w = w16;
        const SplitAct slice{cat16.hi + 256 * j, cat16.lo + 256 * j};
        if ((rc = launch_conv_tc_g(c, g, x, nullptr, nullptr, slice, 1280, B, terms, s))) return rc;
    }
    {
        float* part = (float*)(base + o_gap);                        // [B][GAP_SPLITS][2048], then val [B][256]
        float* val = part + (size_t)B * GAP_SPLITS * 2048;
        enc_gap_kernel<<<dim3(1, B, GAP_SPLITS), 256, 0, s>>>(x.hi, x.lo, use_lo, (int)P16, 2048, part);
        enc_gap_gemv_kernel<<<B, 256, 0, s>>>(part, 1.0f / (float)(h16 * w16), E->gap_w, E->gap_scale, E->gap_shift, val);
        const long long tb = (long long)B * P16 * 32;
        enc_gap_broadcast_kernel<<<(unsigned)((tb + 255) / 256), 256, 0, s>>>(val, cat16.hi, cat16.lo, h16, w16, H16, W16, 1280, 1024, tb);
        c->launches += 3;
        IVOSW_CUDA(cudaGetLastError());
    }
    const SplitAct a16 = view(o_a16, P16, 256);
    {
        GConv g = E->convs[ci++];
        g.in_hp = H16; g.in_wp = W16; g.out_hp = H16; g.out_wp = W16; g.valid_h = h16; g.valid_w = w16;
        if ((rc = launch_conv_tc_g(c, g, cat16, nullptr, nullptr, a16, 256, B, terms, s))) return rc;
    }
    // ---- decoder: [upsampled ASPP | 48-channel shortcut] -> 3x3 -> 3x3 -> depthwise 3x3 -> 1x1 (100)
    const SplitAct cat4 = view(o_cat4, P4, 320);
    {
        const long long tot = (long long)B * P4 * 256 / 8;
        const float sy = h4 > 1 ? (float)(h16 - 1) / (float)(h4 - 1) : 0.f, sx = w4 > 1 ? (float)(w16 - 1) / (float)(w4 - 1) : 0.f;
        enc_upsample_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(a16.hi, a16.lo, use_lo, h16, w16, H16, W16, h4, w4, H4, W4, 256, sy,
                                                                         sx, cat4.hi, cat4.lo, 320, 0, tot);
        c->launches += 1;
        GConv g = E->convs[ci++];                                        // decoder.conv1: 256 -> 48 (64)
        g.in_hp = H4; g.in_wp = W4; g.out_hp = H4; g.out_wp = W4; g.valid_h = h4; g.valid_w = w4;
        const SplitAct slice{cat4.hi + 256, cat4.lo + 256};
        if ((rc = launch_conv_tc_g(c, g, low4, nullptr, nullptr, slice, 320, B, terms, s))) return rc;
    }
    const SplitAct d4a = view(o_d4a, P4, 256), d4b = view(o_d4b, P4, 256), e4 = view(o_e4, P4, 128);
    {
        GConv g = E->convs[ci++];                                        // last_conv.0: 304 (320) -> 256
        g.in_hp = H4; g.in_wp = W4; g.out_hp = H4; g.out_wp = W4; g.valid_h = h4; g.valid_w = w4;
        if ((rc = launch_conv_tc_g(c, g, cat4, nullptr, nullptr, d4a, 256, B, terms, s))) return rc;
        GConv g2 = E->convs[ci++];                                       // last_conv.4
        g2.in_hp = H4; g2.in_wp = W4; g2.out_hp = H4; g2.out_wp = W4; g2.valid_h = h4; g2.valid_w = w4;
        if ((rc = launch_conv_tc_g(c, g2, d4a, nullptr, nullptr, d4b, 256, B, terms, s))) return rc;
        const long long tot = (long long)B * P4 * 256 / 8;
        enc_dwconv_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d4b.hi, d4b.lo, use_lo, E->dw_w, E->dw_scale, E->dw_shift, h4, w4, H4, W4,
                                                                       256, d4a.hi, d4a.lo, tot);
        c->launches += 1;
        GConv g3 = E->convs[ci++];                                       // embed.pw: 256 -> 100 (128)
        g3.in_hp = H4; g3.in_wp = W4; g3.out_hp = H4; g3.out_wp = W4; g3.valid_h = h4; g3.valid_w = w4;
        if ((rc = launch_conv_tc_g(c, g3, d4a, nullptr, nullptr, e4, 128, B, terms, s))) return rc;
        const long long to = (long long)B * 100 * h4 * w4;
        enc_export_kernel<<<(unsigned)((to + 255) / 256), 256, 0, s>>>(e4.hi, e4.lo, use_lo, h4, w4, H4, W4, 128, 100, out, to);
        c->launches += 1;
    }
    IVOSW_CUDA(cudaGetLastError());
    if (ci != E->convs.size()) { set_error("MANet encoder graph walk does not match its layer table"); return IVOSW_ERR_STATE; }
    return IVOSW_OK;
}

}  // namespace ivosw
