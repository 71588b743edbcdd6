// fp32 CUDA-core implicit-GEMM convolution (NHWC activations, KRSC weights) with the eval-mode
// BatchNorm folded to a per-channel scale/shift epilogue, optional residual add and ReLU:
//     out[m, n] = act( (sum_k A[m, k] * W[n, k]) * scale[n] + shift[n] + residual[m, n] )
// m = (b, oh, ow) output pixel, n = output channel, k = (kh, kw, cin).
//
// This is the VALIDATION path (IVOSW_CONV_SIMT_FP32): plain fp32 FMAs, used to pin the tensor-core
// path layer by layer on the device and as the first parity-green implementation.  It serves every
// bottleneck convolution of res2..res5 (torchvision Bottleneck: models/assessment.py:36-39).
#include "ivosw_internal.h"

namespace ivosw {

constexpr int SM_BM = 128, SM_BN = 64, SM_BK = 16;

struct ConvGeom {
    int B, H, W, Cin, OH, OW, Cout, k, stride, pad;
    long long M;  // B * OH * OW
    int K;        // k * k * Cin
};

__global__ void __launch_bounds__(256) conv_simt_kernel(const float* __restrict__ in, const float* __restrict__ wgt,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ shift,
                                                        const float* __restrict__ residual, float* __restrict__ out,
                                                        ConvGeom g, int relu) {
    __shared__ __align__(16) float As[SM_BK][SM_BM + 4];
    __shared__ __align__(16) float Bs[SM_BK][SM_BN + 4];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * SM_BM;
    const int n0 = blockIdx.y * SM_BN;

    // A loader: thread -> rows (tid / 4) and (tid / 4 + 64), 4 consecutive k at (tid % 4) * 4
    const int a_k = (tid & 3) * 4;
    long long a_base[2];   // offset of pixel (b, oh*stride - pad, ow*stride - pad) in `in`, in pixels
    int a_ih0[2], a_iw0[2];
    bool a_ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        long long m = m0 + (tid >> 2) + r * 64;
        a_ok[r] = m < g.M;
        long long mm = a_ok[r] ? m : 0;
        int ow = (int)(mm % g.OW);
        long long t = mm / g.OW;
        int oh = (int)(t % g.OH);
        long long b = t / g.OH;
        a_ih0[r] = oh * g.stride - g.pad;
        a_iw0[r] = ow * g.stride - g.pad;
        a_base[r] = b * g.H * g.W;
    }
    // B loader: thread -> weight row n0 + tid / 4, 4 consecutive k at (tid % 4) * 4
    const float* wrow = wgt + (long long)(n0 + (tid >> 2)) * g.K + a_k;

    const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, each 8 rows x 4 cols
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += SM_BK) {
        // Cin % 16 == 0, so a 16-wide k slab lies inside one filter tap
        const int tap = k0 / g.Cin;
        const int c0 = k0 - tap * g.Cin + a_k;
        const int kh = tap / g.k, kw = tap - kh * g.k;
        float4 av[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int ih = a_ih0[r] + kh, iw = a_iw0[r] + kw;
            av[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_ok[r] && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                av[r] = __ldg(reinterpret_cast<const float4*>(in + (a_base[r] + (long long)ih * g.W + iw) * g.Cin + c0));
        }
        float4 bv = __ldg(reinterpret_cast<const float4*>(wrow + k0));
        __syncthreads();  // previous slab fully consumed
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int row = (tid >> 2) + r * 64;
            As[a_k + 0][row] = av[r].x; As[a_k + 1][row] = av[r].y;
            As[a_k + 2][row] = av[r].z; As[a_k + 3][row] = av[r].w;
        }
        {
            int col = tid >> 2;
            Bs[a_k + 0][col] = bv.x; Bs[a_k + 1][col] = bv.y; Bs[a_k + 2][col] = bv.z; Bs[a_k + 3][col] = bv.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SM_BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    }
    const int n = n0 + tx * 4;
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + n));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + n));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
        float4 o;
        o.x = fmaf(acc[i][0], sc.x, sh.x); o.y = fmaf(acc[i][1], sc.y, sh.y);
        o.z = fmaf(acc[i][2], sc.z, sh.z); o.w = fmaf(acc[i][3], sc.w, sh.w);
        if (residual) {
            float4 rr = __ldg(reinterpret_cast<const float4*>(residual + m * g.Cout + n));
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(out + m * g.Cout + n) = o;
    }
}

// the same kernel with explicit operands (train.cu: raw convolutions, data gradients as convolutions)
int launch_conv_simt_raw(ivosw_ctx* c, const float* in, const float* wgt, const float* scale, const float* shift,
                         const float* residual, float* out, int B, int in_hw, int cin, int out_hw, int cout, int k, int stride,
                         int pad, int relu, cudaStream_t s) {
    if (cin % 16 != 0 || cout % 64 != 0) { set_error("conv_simt: Cin % 16 == 0 and Cout % 64 == 0 required"); return IVOSW_ERR_INVALID; }
    ConvGeom g;
    g.B = B; g.H = in_hw; g.W = in_hw; g.Cin = cin; g.OH = out_hw; g.OW = out_hw; g.Cout = cout;
    g.k = k; g.stride = stride; g.pad = pad;
    g.M = (long long)B * g.OH * g.OW;
    g.K = k * k * cin;
    dim3 grid((unsigned)((g.M + SM_BM - 1) / SM_BM), cout / SM_BN);
    conv_simt_kernel<<<grid, 256, 0, s>>>(in, wgt, scale, shift, residual, out, g, relu);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int launch_conv_simt(ivosw_ctx* c, const ConvLayer& L, const float* in, const float* residual, float* out, int B,
                     cudaStream_t s) {
    ConvGeom g;
    g.B = B; g.H = L.in_hw; g.W = L.in_hw; g.Cin = L.cin; g.OH = L.out_hw; g.OW = L.out_hw; g.Cout = L.cout;
    g.k = L.k; g.stride = L.stride; g.pad = L.pad;
    g.M = (long long)B * g.OH * g.OW;
    g.K = L.k * L.k * L.cin;
    dim3 grid((unsigned)((g.M + SM_BM - 1) / SM_BM), L.cout / SM_BN);
    conv_simt_kernel<<<grid, 256, 0, s>>>(in, L.w_f32, L.scale, L.shift, residual, out, g, L.relu ? 1 : 0);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

}  // namespace ivosw
