// Tail of AssessNet.forward (models/assessment.py:179-180): avg_pool2d(r5, 8) -> fc1 (2048 -> 1),
// and the per-frame glue of recommend_frame (utils/utils_agent.py:120-121): float64 mean over the
// objects and packing of the Brain input state [mask_quality, annotated_count] as fp32.
#include "ivosw_internal.h"

namespace ivosw {

// r5: [B][64][2048] NHWC fp32.  One CTA per sample.
__global__ void __launch_bounds__(256) gap_fc_kernel(const float* __restrict__ r5, const float* __restrict__ fcw,
                                                     float fcb, float* __restrict__ score) {
    const int b = blockIdx.x;
    const float* x = r5 + (long long)b * 64 * 2048;
    float part = 0.f;
    for (int ch = threadIdx.x; ch < 2048; ch += 256) {
        float s = 0.f;
#pragma unroll 8
        for (int p = 0; p < 64; ++p) s += x[p * 2048 + ch];
        part = fmaf(s * (1.0f / 64.0f), __ldg(fcw + ch), part);
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i];
        score[b] = t + fcb;
    }
}

// Same, reading r5 directly in the split-fp16 form the tensor-core path produces (x = hi + lo / 2048):
// no fp32 copy of r5 is materialised.  One CTA per sample, 1024 threads = 4 pixel groups x 256 channel octets,
// 128-bit loads, 16 pixels per thread; the four pixel-group sums are combined in a fixed order (deterministic).
__global__ void __launch_bounds__(1024) gap_fc_split_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                                            int use_lo, const float* __restrict__ fcw, float fcb,
                                                            float* __restrict__ score) {
    const int b = blockIdx.x, oct = threadIdx.x & 255, pg = threadIdx.x >> 8, ch0 = oct * 8;
    const uint4* ph = reinterpret_cast<const uint4*>(hi + ((long long)b * 64 + pg * 16) * 2048 + ch0);
    const uint4* pl = reinterpret_cast<const uint4*>(lo + ((long long)b * 64 + pg * 16) * 2048 + ch0);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int p = 0; p < 16; ++p) {
        const uint4 h4 = __ldg(ph + p * 256);
        const uint4 l4 = use_lo ? __ldg(pl + p * 256) : make_uint4(0, 0, 0, 0);
        const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[u]));
            const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[u]));
            s[2 * u] += fmaf(lf.x, 1.0f / 2048.0f, hf.x);
            s[2 * u + 1] += fmaf(lf.y, 1.0f / 2048.0f, hf.y);
        }
    }
    __shared__ float grp[4][256][8];
    __shared__ float red[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) grp[pg][oct][k] = s[k];
    __syncthreads();
    if (pg == 0) {
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float tot = ((grp[0][oct][k] + grp[1][oct][k]) + grp[2][oct][k]) + grp[3][oct][k];
            part = fmaf(tot * (1.0f / 64.0f), __ldg(fcw + ch0 + k), part);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i];
        score[b] = t + fcb;
    }
}

int launch_gap_fc_split(ivosw_ctx* c, const SplitAct& r5, int use_lo, int B, float* scores, cudaStream_t s) {
    gap_fc_split_kernel<<<B, 1024, 0, s>>>(r5.hi, r5.lo, use_lo, c->fc_w, c->fc_b, scores);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int launch_gap_fc(ivosw_ctx* c, const float* r5, int B, float* scores, cudaStream_t s) {
    gap_fc_kernel<<<B, 256, 0, s>>>(r5, c->fc_w, c->fc_b, scores);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

// scores: [O][T] fp32 (object-major, as the per-object AssessNet calls produce them).
// mq[t] = float64 mean over objects (numpy .mean(1) of a float64 array holding fp32 values);
// state[t] = (float(mq[t]), float(ann[t]))  — torch.Tensor(state[None]) casts to fp32 (agent.py:176).
__global__ void object_mean_kernel(const float* __restrict__ scores, int T, int O, const double* __restrict__ ann,
                                   double* __restrict__ mq, float* __restrict__ state) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    // numpy's pairwise/sequential add for tiny O: plain left-to-right float64 sum, then divide
    double s = 0.0;
    for (int o = 0; o < O; ++o) s += (double)scores[(long long)o * T + t];
    double m = s / (double)O;
    mq[t] = m;
    if (state) {
        state[2 * t + 0] = (float)m;
        state[2 * t + 1] = (float)ann[t];
    }
}

__global__ void pack_state_kernel(const double* __restrict__ mq, const double* __restrict__ ann, int T,
                                  float* __restrict__ state) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    state[2 * t + 0] = (float)mq[t];   // torch.Tensor(state[None]): float64 -> float32 (agent.py:176)
    state[2 * t + 1] = (float)ann[t];
}

int launch_pack_state(ivosw_ctx* c, const double* mq_dev, const double* ann_dev, int T, float* state_dev,
                      cudaStream_t s) {
    pack_state_kernel<<<(T + 127) / 128, 128, 0, s>>>(mq_dev, ann_dev, T, state_dev);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int launch_object_mean(ivosw_ctx* c, const float* scores, int T, int O, const double* ann_dev, double* mq_dev,
                       float* state_dev, cudaStream_t s) {
    object_mean_kernel<<<(T + 127) / 128, 128, 0, s>>>(scores, T, O, ann_dev, mq_dev, state_dev);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

}  // namespace ivosw
