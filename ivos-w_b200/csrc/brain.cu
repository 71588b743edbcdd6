// Q-network: models/agent.py::Brain.forward (lines 33-64) + the first-max argmax of Agent.action
// (lines 187-188), as three launches instead of the reference's ~20 library calls per frame:
//
//   brain_inproj_kernel     all frames in parallel:  e_t = fc2(relu(fc1(x_t)));  gi_t = W_ih e_t.
//                           The same LSTMCell serves both directions (agent.py:48-49), so gi_t is
//                           computed once and shared by the forward and backward chains.
//   brain_recurrent_cluster_kernel
//                           the T-step dependency chain gates = gi_t + W_hh h_{t-1} (zero initial state,
//                           gate order i,f,g,o, no bias: LSTMCell(..., bias=False), agent.py:24-25) on a
//                           cluster of 4 CTAs per (direction, batch row): W_hh (256 KB) lives entirely in
//                           the clusters' registers (each CTA owns 32 hidden units = 128 gate rows, a thread
//                           holds 32 weights), h is exchanged through distributed shared memory once per
//                           step.  brain_recurrent_kernel (one CTA, W_hh half in registers / half in smem)
//                           is the variant that also saves activations for the training step.
//   brain_decode8_kernel    Q_t = fc_d2(relu(fc_d1(relu([h_fw_t ; h_bw_t]))))  (agent.py:55-60), register-blocked
//                           over 8 frames per thread, and argmax over t, first maximum wins (numpy semantics).
//
// Latency-bound (2T dependent steps); weights are 724 KB and stay in L2 / on chip.
#include <cooperative_groups.h>

#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace ivosw {

// offsets (floats) into the canonical parameter blob, see include/ivosw_b200.h
constexpr int P_FC1W = 0;
constexpr int P_FC1B = P_FC1W + 128 * 2;
constexpr int P_FC2W = P_FC1B + 128;
constexpr int P_FC2B = P_FC2W + 128 * 128;
constexpr int P_WIH = P_FC2B + 128;
constexpr int P_WHH = P_WIH + 512 * 128;
constexpr int P_D1W = P_WHH + 512 * 128;
constexpr int P_D1B = P_D1W + 128 * 256;
constexpr int P_D2W = P_D1B + 128;
constexpr int P_D2B = P_D2W + 128;
static_assert(P_D2B + 1 == IVOSW_BRAIN_NUM_PARAMS, "Brain parameter count");

__device__ inline float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) brain_inproj_kernel(const float* __restrict__ P,
                                                           const float* __restrict__ state,  // [N][T][2]
                                                           int T, float* __restrict__ GI,     // [N][T][512]
                                                           float* __restrict__ A1save,        // nullable [N][T][128]
                                                           float* __restrict__ Esave) {       // nullable [N][T][128]
    __shared__ float sa[128], se[128];
    const int t = blockIdx.x, n = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float x0 = state[((long long)n * T + t) * 2 + 0], x1 = state[((long long)n * T + t) * 2 + 1];
    // Every warp owns rows warp, warp + 16, ...: the weight rows of a batch of 8 are fetched before any of them is used
    // (eight independent 512-byte row reads in flight per warp instead of one L2 round trip per row); the per-row
    // arithmetic and its order are unchanged.
    float wv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float* w = P + P_FC2W + (warp + 16 * i) * 128;
        wv[i][0] = w[lane]; wv[i][1] = w[lane + 32]; wv[i][2] = w[lane + 64]; wv[i][3] = w[lane + 96];
    }
    if (threadIdx.x < 128) {
        int j = threadIdx.x;
        float v = fmaf(P[P_FC1W + 2 * j + 1], x1, fmaf(P[P_FC1W + 2 * j], x0, 0.f)) + P[P_FC1B + j];
        sa[j] = fmaxf(v, 0.f);
        if (A1save) A1save[((long long)n * T + t) * 128 + j] = sa[j];
    }
    __syncthreads();
    const float a0 = sa[lane], a1 = sa[lane + 32], a2 = sa[lane + 64], a3 = sa[lane + 96];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = warp + 16 * i;
        float v = wv[i][0] * a0 + wv[i][1] * a1 + wv[i][2] * a2 + wv[i][3] * a3;
        v = warp_sum(v);
        if (lane == 0) {
            se[j] = v + P[P_FC2B + j];   // no ReLU after fc2 (agent.py:46)
            if (Esave) Esave[((long long)n * T + t) * 128 + j] = se[j];
        }
    }
    // first batch of W_ih rows: independent of e, in flight across the barrier
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float* w = P + P_WIH + (warp + 16 * i) * 128;
        wv[i][0] = w[lane]; wv[i][1] = w[lane + 32]; wv[i][2] = w[lane + 64]; wv[i][3] = w[lane + 96];
    }
    __syncthreads();
    const float e0 = se[lane], e1 = se[lane + 32], e2 = se[lane + 64], e3 = se[lane + 96];
    float* gi = GI + ((long long)n * T + t) * 512;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float wn[8][4];
        if (b < 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float* w = P + P_WIH + (warp + 16 * ((b + 1) * 8 + i)) * 128;
                wn[i][0] = w[lane]; wn[i][1] = w[lane + 32]; wn[i][2] = w[lane + 64]; wn[i][3] = w[lane + 96];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = warp + 16 * (b * 8 + i);
            float v = wv[i][0] * e0 + wv[i][1] * e1 + wv[i][2] * e2 + wv[i][3] * e3;
            v = warp_sum(v);
            if (lane == 0) gi[j] = v;
        }
        if (b < 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) wv[i][q] = wn[i][q];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// whh_pack: [32][512] float4, whh_pack[kq][j] = W_hh[j][4kq .. 4kq+3]
__global__ void brain_pack_whh_kernel(const float* __restrict__ whh, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 32 * 512
    if (i >= 32 * 512) return;
    int kq = i / 512, j = i - kq * 512;
    const float* r = whh + j * 128 + kq * 4;
    out[i] = make_float4(r[0], r[1], r[2], r[3]);
}

__device__ inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(512, 1) brain_recurrent_kernel(const float4* __restrict__ whh_pack,
                                                                 const float* __restrict__ GI,  // [N][T][512]
                                                                 int T, float* __restrict__ Hout,   // [N][2][T][128]
                                                                 float* __restrict__ Gsave,         // nullable [N][2][T][512] i,f,g,o
                                                                 float* __restrict__ Csave,         // nullable [N][2][T][128]
                                                                 float* __restrict__ HPsave) {      // nullable [N][2][T][128] h before the step
    extern __shared__ __align__(16) float4 sW[];           // [16][512] float4 : k = 64..127
    float* sg = reinterpret_cast<float*>(sW + 16 * 512);   // 512 gate pre-activations
    float* sh = sg + 512;                                  // 128 hidden state
    const int dir = blockIdx.x, n = blockIdx.y, j = threadIdx.x;
    float4 wr[16];                                         // k = 0..63 in registers
#pragma unroll
    for (int kq = 0; kq < 16; ++kq) wr[kq] = whh_pack[kq * 512 + j];
    for (int kq = 0; kq < 16; ++kq) sW[kq * 512 + j] = whh_pack[(16 + kq) * 512 + j];
    if (j < 128) sh[j] = 0.f;
    float c = 0.f;
    const float* gi = GI + (long long)n * T * 512;
    float* ho = Hout + ((long long)n * 2 + dir) * T * 128;
    int t = dir == 0 ? 0 : T - 1;
    float gi_cur = gi[(long long)t * 512 + j];
    __syncthreads();
    const float4* sh4 = reinterpret_cast<const float4*>(sh);
    for (int s = 0; s < T; ++s) {
        const int tn = dir == 0 ? t + 1 : t - 1;
        float gi_next = 0.f;
        if (s + 1 < T) gi_next = gi[(long long)tn * 512 + j];   // prefetch across the step
        float a0 = gi_cur, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int kq = 0; kq < 16; ++kq) {
            const float4 h = sh4[kq];
            a0 = fmaf(wr[kq].x, h.x, a0); a1 = fmaf(wr[kq].y, h.y, a1);
            a2 = fmaf(wr[kq].z, h.z, a2); a3 = fmaf(wr[kq].w, h.w, a3);
        }
#pragma unroll
        for (int kq = 0; kq < 16; ++kq) {
            const float4 h = sh4[16 + kq];
            const float4 w = sW[kq * 512 + j];
            a0 = fmaf(w.x, h.x, a0); a1 = fmaf(w.y, h.y, a1);
            a2 = fmaf(w.z, h.z, a2); a3 = fmaf(w.w, h.w, a3);
        }
        sg[j] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (j < 128) {
            const float ig = sigmoidf_(sg[j]), fg = sigmoidf_(sg[128 + j]);
            const float gg = tanhf(sg[256 + j]), og = sigmoidf_(sg[384 + j]);
            const long long st = (((long long)n * 2 + dir) * T + t);
            if (Gsave) {
                float* gs = Gsave + st * 512;
                gs[j] = ig; gs[128 + j] = fg; gs[256 + j] = gg; gs[384 + j] = og;
                HPsave[st * 128 + j] = sh[j];
            }
            c = fmaf(fg, c, ig * gg);
            if (Csave) Csave[st * 128 + j] = c;
            const float h = og * tanhf(c);
            sh[j] = h;
            ho[(long long)t * 128 + j] = h;
        }
        __syncthreads();
        t = tn;
        gi_cur = gi_next;
    }
}

// ---------------------------------------------------------------------------------------------
// d1t: decoder_fc1.weight transposed to [256][128] (k-major) so the shared-memory image is a linear copy
__global__ void brain_pack_d1t_kernel(const float* __restrict__ w, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 128 * 256, i = j * 256 + k
    if (i >= 128 * 256) return;
    int j = i >> 8, k = i & 255;
    out[k * 128 + j] = w[i];
}

__global__ void brain_decode8_kernel(const float* __restrict__ P, const float4* __restrict__ d1t,
                                     const float* __restrict__ Hout, int T, float* __restrict__ Q, int* __restrict__ argmax,
                                     unsigned int* __restrict__ done_count);
constexpr int DEC_MAX_SEQ = 65536;      // sequences per launch (one completion counter each)

int brain_pack(ivosw_ctx* c) {
    if (!c->brain_whh_t) IVOSW_CUDA(cudaMalloc(&c->brain_whh_t, sizeof(float4) * 32 * 512));
    brain_pack_whh_kernel<<<(32 * 512 + 255) / 256, 256>>>(c->brain_params + P_WHH, (float4*)c->brain_whh_t);
    if (!c->brain_d1t) IVOSW_CUDA(cudaMalloc(&c->brain_d1t, sizeof(float) * 128 * 256));
    brain_pack_d1t_kernel<<<(128 * 256 + 255) / 256, 256>>>(c->brain_params + P_D1W, c->brain_d1t);
    c->launches += 2;
    IVOSW_CUDA(cudaGetLastError());
    const int rec_smem = 16 * 512 * 16 + (512 + 128) * 4;
    IVOSW_CUDA(cudaFuncSetAttribute(brain_recurrent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rec_smem));
    IVOSW_CUDA(cudaFuncSetAttribute(brain_decode8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (256 * 128 + 256 * 8 + 4 * 8) * 4));
    if (!c->brain_done_count) {
        IVOSW_CUDA(cudaMalloc(&c->brain_done_count, sizeof(unsigned int) * DEC_MAX_SEQ));
        IVOSW_CUDA(cudaMemset(c->brain_done_count, 0, sizeof(unsigned int) * DEC_MAX_SEQ));
    }
    IVOSW_CUDA(cudaDeviceSynchronize());
    return IVOSW_OK;
}

// ---------------------------------------------------------------------------------------------
// Cluster variant of the recurrence (inference path).  grid = (4, 2 directions, N), cluster = 4 CTAs.
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(512, 1)
brain_recurrent_cluster_kernel(const float* __restrict__ P, const float* __restrict__ GI,   // [N][T][512]
                               int T, float* __restrict__ Hout) {                            // [N][2][T][128]
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();          // owns hidden units [32r, 32r + 32)
    const int dir = blockIdx.y, n = blockIdx.z;
    __shared__ __align__(16) float sh[2][128];        // hidden state, double-buffered across steps
    __shared__ float spart[4][4][32];                 // [k quarter][gate][unit]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = warp >> 2, kq = warp & 3;           // this warp: gate g, k in [32kq, 32kq + 32), unit = lane
    float4 w[8];
    {
        const float4* wrow = reinterpret_cast<const float4*>(P + P_WHH + (size_t)(g * 128 + 32 * r + lane) * 128 + 32 * kq);
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __ldg(wrow + i);
    }
    if (threadIdx.x < 128) { sh[0][threadIdx.x] = 0.f; sh[1][threadIdx.x] = 0.f; }
    float c = 0.f;
    const float* gi = GI + (long long)n * T * 512 + 32 * r + lane;
    float* ho = Hout + ((long long)n * 2 + dir) * T * 128 + 32 * r + lane;
    float* peer[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) peer[rr] = cluster.map_shared_rank(&sh[0][0], rr);
    int t = dir == 0 ? 0 : T - 1;
    float gin[4] = {0.f, 0.f, 0.f, 0.f};
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) gin[q] = gi[(long long)t * 512 + q * 128];
    }
    cluster.sync();
    for (int s = 0; s < T; ++s) {
        const int tn = dir == 0 ? t + 1 : t - 1;
        float gnext[4] = {0.f, 0.f, 0.f, 0.f};
        if (warp == 0 && s + 1 < T) {
#pragma unroll
            for (int q = 0; q < 4; ++q) gnext[q] = gi[(long long)tn * 512 + q * 128];   // prefetch across the step
        }
        const float4* h4 = reinterpret_cast<const float4*>(&sh[s & 1][32 * kq]);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 h = h4[i];
            a0 = fmaf(w[i].x, h.x, a0); a1 = fmaf(w[i].y, h.y, a1);
            a2 = fmaf(w[i].z, h.z, a2); a3 = fmaf(w[i].w, h.w, a3);
        }
        spart[kq][g][lane] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (warp == 0) {
            float pre[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                pre[q] = gin[q] + ((spart[0][q][lane] + spart[1][q][lane]) + (spart[2][q][lane] + spart[3][q][lane]));
            const float ig = sigmoidf_(pre[0]), fg = sigmoidf_(pre[1]), gg = tanhf(pre[2]), og = sigmoidf_(pre[3]);
            c = fmaf(fg, c, ig * gg);
            const float h = og * tanhf(c);
            ho[(long long)t * 128] = h;
            const int slot = ((s + 1) & 1) * 128 + 32 * r + lane;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) peer[rr][slot] = h;      // distributed shared memory broadcast
#pragma unroll
            for (int q = 0; q < 4; ++q) gin[q] = gnext[q];
        }
        cluster.sync();                                             // h visible cluster-wide; spart reusable
        t = tn;
    }
}

// Decoder + argmax.  One CTA of 128 threads (= the 128 decoder units) per group of 8 frames: every weight read from
// shared memory feeds 8 FMAs, the 128 KB weight image arrives as four bulk copies (one elected thread, overlapped with
// the loads of the hidden states), and the groups of a sequence run on different SMs instead of queueing on one CTA's
// shared-memory pipe.  The CTA that finishes last (one atomic counter per sequence) takes the arg-max over t, first
// maximum wins (numpy semantics).  Per-output arithmetic and its order are those of the single-CTA version.
constexpr int DEC_SMEM_BYTES = (256 * 128 + 256 * 8 + 4 * 8) * 4;
__global__ void __launch_bounds__(128, 1) brain_decode8_kernel(const float* __restrict__ P,
                                                               const float4* __restrict__ d1t,
                                                               const float* __restrict__ Hout,  // [N][2][T][128]
                                                               int T, float* __restrict__ Q,     // [N][T]
                                                               int* __restrict__ argmax,
                                                               unsigned int* __restrict__ done_count) {   // [N], zero between launches
    extern __shared__ __align__(128) float sm[];
    float* sWt = sm;                              // [256][128]
    float* ss = sWt + 256 * 128;                  // [256][8 frames]
    float* sred = ss + 256 * 8;                   // [4 warps][8 frames]
    __shared__ uint64_t wbar;
    __shared__ int is_last;
    const int grp = blockIdx.x, n = blockIdx.y, j = threadIdx.x, wig = j >> 5, lane = j & 31;
    if (j == 0) { mbar_init(&wbar, 1); fence_barrier_init(); }
    __syncthreads();
    if (j == 0) {
        mbar_expect_tx(&wbar, 256 * 128 * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sWt + i * 8192)), "l"(reinterpret_cast<const float*>(d1t) + i * 8192), "r"(32768),
                           "r"(smem_u32(&wbar)) : "memory");
    }
    const float b1 = P[P_D1B + j], w2 = P[P_D2W + j], b2 = P[P_D2B];
    const float* hf = Hout + ((long long)n * 2 + 0) * T * 128;
    const float* hb = Hout + ((long long)n * 2 + 1) * T * 128;
    const int tb = grp * 8;
    float vf[8], vb[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        const int t = tb + f;
        vf[f] = t < T ? fmaxf(hf[(long long)t * 128 + j], 0.f) : 0.f;
        vb[f] = t < T ? fmaxf(hb[(long long)t * 128 + j], 0.f) : 0.f;
    }
    reinterpret_cast<float4*>(ss + j * 8)[0] = make_float4(vf[0], vf[1], vf[2], vf[3]);
    reinterpret_cast<float4*>(ss + j * 8)[1] = make_float4(vf[4], vf[5], vf[6], vf[7]);
    reinterpret_cast<float4*>(ss + (128 + j) * 8)[0] = make_float4(vb[0], vb[1], vb[2], vb[3]);
    reinterpret_cast<float4*>(ss + (128 + j) * 8)[1] = make_float4(vb[4], vb[5], vb[6], vb[7]);
    __syncthreads();
    mbar_wait(&wbar, 0);                          // weight image landed
    float acc[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) acc[f] = b1;
#pragma unroll 4
    for (int k = 0; k < 256; ++k) {
        const float wv = sWt[k * 128 + j];
        const float4 s0 = reinterpret_cast<const float4*>(ss + k * 8)[0];
        const float4 s1 = reinterpret_cast<const float4*>(ss + k * 8)[1];
        acc[0] = fmaf(wv, s0.x, acc[0]); acc[1] = fmaf(wv, s0.y, acc[1]);
        acc[2] = fmaf(wv, s0.z, acc[2]); acc[3] = fmaf(wv, s0.w, acc[3]);
        acc[4] = fmaf(wv, s1.x, acc[4]); acc[5] = fmaf(wv, s1.y, acc[5]);
        acc[6] = fmaf(wv, s1.z, acc[6]); acc[7] = fmaf(wv, s1.w, acc[7]);
    }
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        const float part = warp_sum(w2 * fmaxf(acc[f], 0.f));
        if (lane == 0) sred[wig * 8 + f] = part;
    }
    __syncthreads();
    if (j < 8 && tb + j < T) {
        const float* rp = sred + j;
        Q[(long long)n * T + tb + j] = ((rp[0] + rp[8]) + (rp[16] + rp[24])) + b2;
    }
    if (argmax == nullptr) return;
    __threadfence();                              // this group's Q values before the counter
    __syncthreads();
    if (j == 0) is_last = atomicAdd(done_count + n, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last && j < 32) {   // first maximum wins (numpy argmax)
        __threadfence();
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int t = lane; t < T; t += 32) {
            float v = __ldcg(Q + (long long)n * T + t);
            if (v > best) { best = v; bi = t; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { argmax[n] = (bi == 0x7fffffff) ? 0 : bi; done_count[n] = 0u; }
    }
}

// params / whh_pack / d1t select the network (policy or target); the save pointers are for the training step
int launch_brain_ex(ivosw_ctx* c, const float* params, const float* whh_pack, const float* d1t, const float* state,
                    int N, int T, float* q, int* argmax, const BrainSaves* sv, cudaStream_t s) {
    int rc;
    if ((rc = ensure(c->brain_gi, sizeof(float) * (size_t)N * T * 512))) return rc;
    if ((rc = ensure(c->brain_h, sizeof(float) * (size_t)N * 2 * T * 128))) return rc;
    float* hout = sv && sv->H ? sv->H : (float*)c->brain_h.p;
    brain_inproj_kernel<<<dim3(T, N), 512, 0, s>>>(params, state, T, (float*)c->brain_gi.p, sv ? sv->A1 : nullptr,
                                                   sv ? sv->E : nullptr);
    IVOSW_CUDA(cudaGetLastError());
    if (sv == nullptr) {     // inference: cluster recurrence, W_hh in registers
        brain_recurrent_cluster_kernel<<<dim3(4, 2, N), 512, 0, s>>>(params, (const float*)c->brain_gi.p, T, hout);
    } else {                 // training step: single-CTA recurrence that also saves gates / cell states
        const int rec_smem = 16 * 512 * 16 + (512 + 128) * 4;
        brain_recurrent_kernel<<<dim3(2, N), 512, rec_smem, s>>>((const float4*)whh_pack, (const float*)c->brain_gi.p, T, hout,
                                                                 sv->G, sv->C, sv->HP);
    }
    IVOSW_CUDA(cudaGetLastError());
    if (N > DEC_MAX_SEQ) { set_error("Brain: more than 65536 sequences per call"); return IVOSW_ERR_INVALID; }
    if (!c->brain_done_count) { set_error("Brain weights not loaded"); return IVOSW_ERR_STATE; }
    brain_decode8_kernel<<<dim3((T + 7) / 8, N), 128, DEC_SMEM_BYTES, s>>>(params, (const float4*)d1t, hout, T, q, argmax,
                                                                          c->brain_done_count);
    IVOSW_CUDA(cudaGetLastError());
    c->launches += 3;
    return IVOSW_OK;
}

int launch_brain(ivosw_ctx* c, const float* state, int N, int T, float* q, int* argmax, cudaStream_t s) {
    return launch_brain_ex(c, c->brain_params, c->brain_whh_t, c->brain_d1t, state, N, T, q, argmax, nullptr, s);
}

// packs W_hh / decoder_fc1 of an arbitrary parameter blob (used for the target network and after each update)
int brain_pack_into(ivosw_ctx* c, const float* params, float* whh_pack, float* d1t, cudaStream_t s) {
    brain_pack_whh_kernel<<<(32 * 512 + 255) / 256, 256, 0, s>>>(params + P_WHH, (float4*)whh_pack);
    brain_pack_d1t_kernel<<<(128 * 256 + 255) / 256, 256, 0, s>>>(params + P_D1W, d1t);
    c->launches += 2;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

}  // namespace ivosw
