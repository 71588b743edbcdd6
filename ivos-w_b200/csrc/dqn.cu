// Double-DQN training step of the Q-network: models/agent.py::Agent.update_agent (lines 103-166),
// BASELINE config C4 (batch 256 x T 25).  Everything stays on the device:
//
//   no-grad:  a* = argmax_t Q_policy(s'),  Q_next = Q_target(s')[a*]                     (:131-139)
//             y_step = gamma * Q_next + 0.1 * r_step,   y_done = 0.1 * r_done             (:140-141)
//   forward with saved activations:  q = Q_policy(s)[a]                                   (:146-147)
//   loss = mse(q, y_step) + mse(q, y_done)                                                (:151-153)
//   backward: the loss touches ONE frame per sample, so each sample's backward is a decoder step at
//             t = a followed by two LSTM chains (forward cell: t = a..0, backward cell: t = a..T-1);
//             weight gradients are deterministic GEMM-shaped reductions over the saved activations
//   element-wise gradient clamp to +-1 (:157-159), Adam with L2 weight decay in the gradient (:101,160).
// The stochastic hard target sync (:163-165) is the caller's decision (ivosw_dqn_sync_target).
#include <cmath>
#include <cstdio>

#include "ivosw_internal.h"

namespace ivosw {

namespace {

constexpr int P_FC1W = 0, P_FC1B = 256, P_FC2W = 384, P_FC2B = 384 + 16384, P_WIH = P_FC2B + 128;
constexpr int P_WHH = P_WIH + 65536, P_D1W = P_WHH + 65536, P_D1B = P_D1W + 32768, P_D2W = P_D1B + 128;
constexpr int P_D2B = P_D2W + 128;
constexpr int NPARAM = IVOSW_BRAIN_NUM_PARAMS;

// ---- decoder at t = action[n]: targets, loss terms, dq, gradients w.r.t. the two hidden states ----
__global__ void __launch_bounds__(256) dec_bwd_kernel(const float* __restrict__ P, const float* __restrict__ H,
                                                      const int* __restrict__ action, const float* __restrict__ q_tgt_new,
                                                      const int* __restrict__ a_star, const float* __restrict__ r_step,
                                                      const float* __restrict__ r_done, int N, int T, float gamma,
                                                      float* __restrict__ S,      // [N][256] relu'd state
                                                      float* __restrict__ DD,     // [N][128] grad at decoder_fc1 pre-activation
                                                      float* __restrict__ Dact,   // [N][128] relu(fc_d1)
                                                      float* __restrict__ DQ,     // [N]
                                                      float* __restrict__ DH,     // [N][2][128] grad w.r.t. h_fw(a), h_bw(a)
                                                      float* __restrict__ loss_part) {
    __shared__ float s[256], dd[128], red[8];
    const int n = blockIdx.x, tid = threadIdx.x, a = action[n];
    const float* hf = H + (((long long)n * 2 + 0) * T + a) * 128;
    const float* hb = H + (((long long)n * 2 + 1) * T + a) * 128;
    const float raw = tid < 128 ? hf[tid] : hb[tid - 128];
    s[tid] = fmaxf(raw, 0.f);
    S[(long long)n * 256 + tid] = s[tid];
    __syncthreads();
    float dj = 0.f, zj = 0.f;
    if (tid < 128) {
        const float* w = P + P_D1W + tid * 256;
        float a0 = P[P_D1B + tid], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
        for (int k = 0; k < 256; k += 4) {
            a0 = fmaf(w[k], s[k], a0); a1 = fmaf(w[k + 1], s[k + 1], a1);
            a2 = fmaf(w[k + 2], s[k + 2], a2); a3 = fmaf(w[k + 3], s[k + 3], a3);
        }
        zj = (a0 + a1) + (a2 + a3);
        dj = fmaxf(zj, 0.f);
        Dact[(long long)n * 128 + tid] = dj;
    }
    // q = wd2 . d + bd2
    float part = tid < 128 ? P[P_D2W + tid] * dj : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    const float q = ((red[0] + red[1]) + (red[2] + red[3])) + P[P_D2B];
    const float y1 = fmaf(q_tgt_new[(long long)n * T + a_star[n]], gamma, r_step[n] * 0.1f);   // :140
    const float y2 = r_done[n] * 0.1f;                                                             // :141
    const float dq = (2.0f / (float)N) * ((q - y1) + (q - y2));
    if (tid == 0) {
        DQ[n] = dq;
        loss_part[n] = (q - y1) * (q - y1) + (q - y2) * (q - y2);
    }
    if (tid < 128) {
        const float g = zj > 0.f ? dq * P[P_D2W + tid] : 0.f;
        dd[tid] = g;
        DD[(long long)n * 128 + tid] = g;
    }
    __syncthreads();
    // ds_k = sum_j Wd1[j][k] dd_j, masked by the ReLU on the concatenated state
    float acc = 0.f;
    for (int j = 0; j < 128; ++j) acc = fmaf(P[P_D1W + j * 256 + tid], dd[j], acc);
    DH[(long long)n * 256 + tid] = raw > 0.f ? acc : 0.f;
}

// ---- LSTM backward chain of one (direction, sample) ----
__global__ void __launch_bounds__(512) lstm_bwd_kernel(const float* __restrict__ P, const float* __restrict__ G,
                                                       const float* __restrict__ C, const int* __restrict__ action,
                                                       const float* __restrict__ DH, int T,
                                                       float* __restrict__ DG) {   // [N][2][T][512], pre-zeroed
    __shared__ float sdg[512], sdh[128], part[4][128];
    const int dir = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
    const int a = action[n];
    const long long base = ((long long)n * 2 + dir) * T;
    float dc = 0.f;
    if (tid < 128) sdh[tid] = DH[((long long)n * 2 + dir) * 128 + tid];
    __syncthreads();
    const int steps = dir == 0 ? a + 1 : T - a;
    for (int s = 0; s < steps; ++s) {
        const int t = dir == 0 ? a - s : a + s;
        const int tprev = dir == 0 ? t - 1 : t + 1;           // the cell state this step started from
        const bool has_prev = dir == 0 ? (t > 0) : (t < T - 1);
        if (tid < 128) {
            const float* g = G + (base + t) * 512;
            const float ig = g[tid], fg = g[128 + tid], gg = g[256 + tid], og = g[384 + tid];
            const float ct = C[(base + t) * 128 + tid];
            const float cp = has_prev ? C[(base + tprev) * 128 + tid] : 0.f;
            const float tc = tanhf(ct);
            const float dh = sdh[tid];
            const float dct = dc + dh * og * (1.f - tc * tc);
            const float d_o = dh * tc, d_i = dct * gg, d_g = dct * ig, d_f = dct * cp;
            dc = dct * fg;
            sdg[tid] = d_i * ig * (1.f - ig);
            sdg[128 + tid] = d_f * fg * (1.f - fg);
            sdg[256 + tid] = d_g * (1.f - gg * gg);
            sdg[384 + tid] = d_o * og * (1.f - og);
        }
        __syncthreads();
        DG[(base + t) * 512 + tid] = sdg[tid];
        // dh_prev[k] = sum_j W_hh[j][k] dgate_j : 4 partial sums of 128 j each, coalesced over k
        {
            const int k = tid & 127, pj = tid >> 7;
            const float* w = P + P_WHH + (pj * 128) * 128 + k;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
            for (int j = 0; j < 128; j += 2) {
                a0 = fmaf(w[j * 128], sdg[pj * 128 + j], a0);
                a1 = fmaf(w[(j + 1) * 128], sdg[pj * 128 + j + 1], a1);
            }
            part[pj][k] = a0 + a1;
        }
        __syncthreads();
        if (tid < 128) sdh[tid] = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
        __syncthreads();
    }
}

// ---- generic fp32 GEMM, C[M][N] = op(A) * op(B), deterministic; TA: A stored [K][M] (use A^T); TB: B stored [N][K]
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                   float* __restrict__ Cm, int M, int N, int K, int lda, int ldb, int ldc) {
    __shared__ float As[16][64 + 1], Bs[16][64 + 1];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64, tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = tid; i < 16 * 64; i += 256) {
            const int kk = i >> 6, mm = i & 63;
            const int gm = m0 + mm, gk = k0 + kk;
            float v = 0.f;
            if (gm < M && gk < K) v = TA ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk];
            As[kk][mm] = v;
            const int gn = n0 + mm;
            float u = 0.f;
            if (gn < N && gk < K) u = TB ? B[(long long)gn * ldb + gk] : B[(long long)gk * ldb + gn];
            Bs[kk][mm] = u;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
            if (gm < M && gn < N) Cm[(long long)gm * ldc + gn] = acc[i][j];
        }
}

template <bool TA, bool TB>
int gemm(ivosw_ctx* c, const float* A, const float* B, float* Cm, int M, int N, int K, int lda, int ldb, int ldc,
         cudaStream_t s) {
    gemm_kernel<TA, TB><<<dim3((N + 63) / 64, (M + 63) / 64), 256, 0, s>>>(A, B, Cm, M, N, K, lda, ldb, ldc);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

// out[j] = sum_r A[r][j]  (deterministic: one thread per column, sequential over rows in 4 interleaved partials)
__global__ void colsum_kernel(const float* __restrict__ A, int R, int J, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int r = 0;
    for (; r + 3 < R; r += 4) {
        a0 += A[(long long)r * J + j]; a1 += A[(long long)(r + 1) * J + j];
        a2 += A[(long long)(r + 2) * J + j]; a3 += A[(long long)(r + 3) * J + j];
    }
    for (; r < R; ++r) a0 += A[(long long)r * J + j];
    out[j] = (a0 + a1) + (a2 + a3);
}

// DGs[n][t][:] = DG[n][0][t][:] + DG[n][1][t][:]
__global__ void sum_dirs_kernel(const float* __restrict__ DG, int T, long long total, float* __restrict__ DGs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over N*T*512
    if (i >= total) return;
    const int j = (int)(i & 511);
    const long long nt = i >> 9;
    const long long n = nt / T, t = nt - n * T;
    DGs[i] = DG[((n * 2 + 0) * T + t) * 512 + j] + DG[((n * 2 + 1) * T + t) * 512 + j];
}

__global__ void relu_mask_kernel(float* __restrict__ d, const float* __restrict__ act, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(act[i] > 0.f)) d[i] = 0.f;
}

// action[n] indexes the frame axis of the saved activations: a replay sample outside [0, T) would read (and, in the LSTM
// chain, write) out of bounds.  The reference's gather() raises on such an index; so does this path, before touching anything.
__global__ void check_actions_kernel(const int* __restrict__ action, int N, int T, int* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && (action[i] < 0 || action[i] >= T)) atomicExch(bad, i + 1);
}

__global__ void loss_reduce_kernel(const float* __restrict__ part, int N, float* __restrict__ loss) {
    __shared__ float red[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) a += part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        *loss = t / (float)N;
    }
}

// clamp -> (+ wd * p) -> Adam, exactly torch.optim.Adam's update order (agent.py:157-160)
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int n, float lr, float wd, float bc1, float bc2_sqrt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = fminf(fmaxf(g[i], -1.f), 1.f);
    g[i] = gi;                                   // the clamped gradient is what tests compare
    gi = fmaf(wd, p[i], gi);
    const float mi = 0.9f * m[i] + (1.f - 0.9f) * gi;
    const float vi = 0.999f * v[i] + (1.f - 0.999f) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + 1e-8f;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
}

}  // namespace

// clamp + Adam(weight decay) on the policy parameters, in place; `grad` is overwritten with the clamped gradient
int dqn_apply(ivosw_ctx* c, float* grad, float lr, float weight_decay, cudaStream_t s) {
    int rc;
    if (!c->adam_m) {
        IVOSW_CUDA(cudaMalloc(&c->adam_m, sizeof(float) * NPARAM));
        IVOSW_CUDA(cudaMalloc(&c->adam_v, sizeof(float) * NPARAM));
        IVOSW_CUDA(cudaMemsetAsync(c->adam_m, 0, sizeof(float) * NPARAM, s));
        IVOSW_CUDA(cudaMemsetAsync(c->adam_v, 0, sizeof(float) * NPARAM, s));
        c->adam_step = 0;
    }
    c->adam_step += 1;
    const float bc1 = (float)(1.0 - pow(0.9, (double)c->adam_step));            // torch computes these in double
    const float bc2s = (float)sqrt(1.0 - pow(0.999, (double)c->adam_step));
    adam_kernel<<<(NPARAM + 255) / 256, 256, 0, s>>>(c->brain_params, grad, c->adam_m, c->adam_v, NPARAM, lr, weight_decay,
                                                     bc1, bc2s);
    IVOSW_CUDA(cudaGetLastError());
    c->launches += 1;
    if ((rc = brain_pack_into(c, c->brain_params, c->brain_whh_t, c->brain_d1t, s))) return rc;   // inference copies
    return IVOSW_OK;
}

int dqn_update(ivosw_ctx* c, const float* state, const float* new_state, const int* action, const float* reward_step,
               const float* reward_done, int N, int T, float gamma, float lr, float weight_decay, float* loss_host,
               float* grads_out_dev, bool apply, cudaStream_t s) {
    int rc;
    if (!c->brain_loaded || !c->target_loaded) { set_error("policy / target Brain weights not loaded"); return IVOSW_ERR_STATE; }
    const size_t NT = (size_t)N * T;
    // workspace carve-up (floats)
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 63) & ~(size_t)63; return o; };
    const size_t o_qpn = take(NT), o_qtn = take(NT), o_q = take(NT), o_astar = take(N), o_a1 = take(NT * 128),
                 o_e = take(NT * 128), o_g = take(NT * 2 * 512), o_c = take(NT * 2 * 128), o_hp = take(NT * 2 * 128),
                 o_h = take(NT * 2 * 128), o_s = take((size_t)N * 256), o_dd = take((size_t)N * 128),
                 o_dact = take((size_t)N * 128), o_dq = take(N), o_dh = take((size_t)N * 256), o_lp = take(N),
                 o_dg = take(NT * 2 * 512), o_dgs = take(NT * 512), o_de = take(NT * 128), o_da1 = take(NT * 128),
                 o_grad = take(NPARAM), o_loss = take(64);
    if ((rc = ensure(c->dqn_ws, off * sizeof(float)))) return rc;
    float* W = (float*)c->dqn_ws.p;
    float* grad = W + o_grad;
    const float* Pp = c->brain_params;
    {   // validate the replay actions first (one tiny kernel + a 4-byte read; the step itself takes milliseconds)
        if ((rc = ensure_pinned(c, 64))) return rc;
        int* bad_dev = (int*)(W + o_loss) + 8;
        IVOSW_CUDA(cudaMemsetAsync(bad_dev, 0, sizeof(int), s));
        check_actions_kernel<<<(N + 255) / 256, 256, 0, s>>>(action, N, T, bad_dev);
        IVOSW_CUDA(cudaGetLastError());
        c->launches += 1;
        IVOSW_CUDA(cudaMemcpyAsync(c->pinned_small, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
        IVOSW_CUDA(cudaStreamSynchronize(s));
        const int bad = *(int*)c->pinned_small;
        if (bad) {
            char buf[160];
            snprintf(buf, sizeof buf, "invalid argument: action[%d] is outside [0, T=%d) (index out of range in gather)", bad - 1, T);
            set_error(buf);
            return IVOSW_ERR_INVALID;
        }
    }
    // ---- no-grad forwards on the new state (policy -> a*, target -> Q_next)
    if ((rc = launch_brain_ex(c, Pp, c->brain_whh_t, c->brain_d1t, new_state, N, T, W + o_qpn, (int*)(W + o_astar), nullptr, s)))
        return rc;
    if ((rc = launch_brain_ex(c, c->target_params, c->target_whh_t, c->target_d1t, new_state, N, T, W + o_qtn, nullptr,
                              nullptr, s)))
        return rc;
    // ---- forward on the state, activations kept
    BrainSaves sv{W + o_a1, W + o_e, W + o_g, W + o_c, W + o_hp, W + o_h};
    if ((rc = launch_brain_ex(c, Pp, c->brain_whh_t, c->brain_d1t, state, N, T, W + o_q, nullptr, &sv, s))) return rc;
    // ---- backward
    IVOSW_CUDA(cudaMemsetAsync(W + o_dg, 0, sizeof(float) * NT * 2 * 512, s));
    dec_bwd_kernel<<<N, 256, 0, s>>>(Pp, W + o_h, action, W + o_qtn, (const int*)(W + o_astar), reward_step, reward_done,
                                     N, T, gamma, W + o_s, W + o_dd, W + o_dact, W + o_dq, W + o_dh, W + o_lp);
    IVOSW_CUDA(cudaGetLastError());
    lstm_bwd_kernel<<<dim3(2, N), 512, 0, s>>>(Pp, W + o_g, W + o_c, action, W + o_dh, T, W + o_dg);
    IVOSW_CUDA(cudaGetLastError());
    loss_reduce_kernel<<<1, 256, 0, s>>>(W + o_lp, N, W + o_loss);
    const long long tot = (long long)NT * 512;
    sum_dirs_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(W + o_dg, T, tot, W + o_dgs);
    IVOSW_CUDA(cudaGetLastError());
    c->launches += 4;
    const int R2 = (int)(NT * 2), R1 = (int)NT;
    // decoder gradients
    if ((rc = gemm<true, false>(c, W + o_dd, W + o_s, grad + P_D1W, 128, 256, N, 128, 256, 256, s))) return rc;     // dWd1 = DD^T S
    colsum_kernel<<<1, 128, 0, s>>>(W + o_dd, N, 128, grad + P_D1B);
    if ((rc = gemm<true, false>(c, W + o_dq, W + o_dact, grad + P_D2W, 1, 128, N, 1, 128, 128, s))) return rc;       // dwd2 = DQ^T D
    colsum_kernel<<<1, 32, 0, s>>>(W + o_dq, N, 1, grad + P_D2B);
    // LSTM weight gradients
    if ((rc = gemm<true, false>(c, W + o_dg, W + o_hp, grad + P_WHH, 512, 128, R2, 512, 128, 128, s))) return rc;    // dWhh = DG^T Hprev
    if ((rc = gemm<true, false>(c, W + o_dgs, W + o_e, grad + P_WIH, 512, 128, R1, 512, 128, 128, s))) return rc;    // dWih = DGs^T E
    // encoder
    if ((rc = gemm<false, false>(c, W + o_dgs, Pp + P_WIH, W + o_de, R1, 128, 512, 512, 128, 128, s))) return rc;    // DE = DGs Wih
    if ((rc = gemm<true, false>(c, W + o_de, W + o_a1, grad + P_FC2W, 128, 128, R1, 128, 128, 128, s))) return rc;   // dW2 = DE^T A1
    colsum_kernel<<<1, 128, 0, s>>>(W + o_de, R1, 128, grad + P_FC2B);
    if ((rc = gemm<false, false>(c, W + o_de, Pp + P_FC2W, W + o_da1, R1, 128, 128, 128, 128, 128, s))) return rc;   // DA1 = DE W2
    relu_mask_kernel<<<(unsigned)((NT * 128 + 255) / 256), 256, 0, s>>>(W + o_da1, W + o_a1, (long long)NT * 128);
    if ((rc = gemm<true, false>(c, W + o_da1, state, grad + P_FC1W, 128, 2, R1, 128, 2, 2, s))) return rc;           // dW1 = DA1^T X
    colsum_kernel<<<1, 128, 0, s>>>(W + o_da1, R1, 128, grad + P_FC1B);
    IVOSW_CUDA(cudaGetLastError());
    c->launches += 5;
    if (grads_out_dev)   // raw (unclamped) gradients when the optimiser step is deferred, clamped ones otherwise
        if (!apply) IVOSW_CUDA(cudaMemcpyAsync(grads_out_dev, grad, sizeof(float) * NPARAM, cudaMemcpyDeviceToDevice, s));
    if (apply) {
        if ((rc = dqn_apply(c, grad, lr, weight_decay, s))) return rc;
        if (grads_out_dev)
            IVOSW_CUDA(cudaMemcpyAsync(grads_out_dev, grad, sizeof(float) * NPARAM, cudaMemcpyDeviceToDevice, s));
    }
    if ((rc = ensure_pinned(c, 64))) return rc;
    IVOSW_CUDA(cudaMemcpyAsync(c->pinned_small, W + o_loss, sizeof(float), cudaMemcpyDeviceToHost, s));
    IVOSW_CUDA(cudaStreamSynchronize(s));
    if (loss_host) *loss_host = *(float*)c->pinned_small;
    return IVOSW_OK;
}

}  // namespace ivosw
