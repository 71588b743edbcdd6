// The one exchange step of the frame-sharded round: every rank needs the float64 quality of ALL frames
// before the bidirectional LSTM of the Q-network can run.  The message is tiny (ceil(T / G) doubles per rank: 64 bytes at
// T = 64, G = 8), so the cost of an NCCL all-gather is pure launch + protocol latency on the critical path of a 1.3 ms
// step.  Here the exchange is two small kernels of this library over NVLink peer memory instead:
//
//   gather_post_kernel   after a rank has scored its frames: stores its slice straight into EVERY rank's gather buffer
//                        (peer stores through NVLink / NVSwitch, cudaIpc-mapped), __threadfence_system, then raises its flag
//                        in every rank's buffer (st.release.sys of the round's epoch number)
//   gather_wait_pack_kernel   before Brain: waits (ld.acquire.sys) until all G flags of this rank's own buffer carry the
//                        current epoch, then packs the (quality, annotated-count) state for the Q-network
//
// No host synchronisation, no NCCL call, both kernels are ordinary stream work (CUDA-graph capturable: the epoch lives in
// device memory and is advanced by gather_post_kernel itself, so no kernel argument changes from round to round).
// Buffers are double-buffered by epoch parity: a fast rank can be at most one round ahead of the slowest one (its next
// round cannot finish before every rank has posted that round), so round k + 1 never overwrites values of round k that a
// slow rank has not read yet.
//
// Buffer of one rank (device memory, cudaMalloc, exported with cudaIpcGetMemHandle):
//   double            mq[2][cap]        gathered quality values, slot = epoch & 1, rank r's slice at r * per
//   unsigned long long flag[2][64]      flag[slot][r] = epoch once rank r's slice of that epoch is complete
//   unsigned long long epoch            this rank's round counter (only ever touched by this rank)
#include <cstring>

#include "ivosw_internal.h"

namespace ivosw {

constexpr int GATHER_MAX_WORLD = 64;

struct GatherPeers {
    double* mq[GATHER_MAX_WORLD];
    unsigned long long* flag[GATHER_MAX_WORLD];
};

struct GatherState {
    int world = 0, rank = 0, cap = 0;
    void* base = nullptr;                       // this rank's buffer
    void* peer_base[GATHER_MAX_WORLD] = {};     // every rank's buffer as mapped into this process (own: base)
    bool opened = false;
    GatherPeers peers;
    unsigned long long* epoch = nullptr;
};

static size_t gather_bytes(int cap) { return sizeof(double) * 2 * (size_t)cap + sizeof(unsigned long long) * (2 * 64 + 8); }

__global__ void __launch_bounds__(256) gather_post_kernel(GatherPeers P, unsigned long long* epoch, const double* __restrict__ mq_local,
                                                          int n_local, int offset, int cap, int world, int rank) {
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) { s_epoch = *epoch + 1; *epoch = s_epoch; }
    __syncthreads();
    const unsigned long long e = s_epoch;
    const int slot = (int)(e & 1ull);
    for (int i = threadIdx.x; i < n_local * world; i += blockDim.x) {
        const int p = i / n_local, t = i - p * n_local;
        P.mq[p][(size_t)slot * cap + offset + t] = mq_local[t];              // peer store (own buffer for p == rank)
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned long long* f = P.flag[threadIdx.x] + slot * 64 + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(e) : "memory");
    }
}

__global__ void __launch_bounds__(256) gather_wait_pack_kernel(const double* __restrict__ mq_base, const unsigned long long* flag_base,
                                                               const unsigned long long* epoch, int cap, int world,
                                                               const double* __restrict__ ann, int T, float* __restrict__ state,
                                                               double* __restrict__ mq_out) {
    const unsigned long long e = *epoch;                                     // advanced by this round's gather_post_kernel
    const int slot = (int)(e & 1ull);
    if (threadIdx.x < world) {
        const unsigned long long* f = flag_base + slot * 64 + threadIdx.x;
        long long t0 = 0;
        for (unsigned it = 0;; ++it) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= e) break;
            if ((it & 255u) == 255u) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ll) __trap();                  // ~2 s: a rank never posted
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    const double* mq = mq_base + (size_t)slot * cap;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const double m = mq[t];
        state[2 * t + 0] = (float)m;                                         // torch.Tensor(state[None]): float64 -> float32
        state[2 * t + 1] = (float)ann[t];
        if (mq_out) mq_out[t] = m;
    }
}

void gather_release(ivosw_ctx* c) {
    GatherState* G = static_cast<GatherState*>(c->gather_state);
    if (!G) return;
    if (G->opened)
        for (int r = 0; r < G->world; ++r)
            if (r != G->rank && G->peer_base[r]) cudaIpcCloseMemHandle(G->peer_base[r]);
    if (G->base) cudaFree(G->base);
    delete G;
    c->gather_state = nullptr;
}

int gather_create(ivosw_ctx* c, int world, int rank, int cap, void* handle_out) {
    gather_release(c);
    GatherState* G = new GatherState();
    c->gather_state = G;
    G->world = world; G->rank = rank; G->cap = cap;
    IVOSW_CUDA(cudaMalloc(&G->base, gather_bytes(cap)));
    IVOSW_CUDA(cudaMemset(G->base, 0, gather_bytes(cap)));
    cudaIpcMemHandle_t h;
    IVOSW_CUDA(cudaIpcGetMemHandle(&h, G->base));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    memcpy(handle_out, &h, sizeof h);
    return IVOSW_OK;
}

int gather_open(ivosw_ctx* c, const void* handles) {
    GatherState* G = static_cast<GatherState*>(c->gather_state);
    if (!G) { set_error("ivosw_gather_create has not been called"); return IVOSW_ERR_STATE; }
    for (int r = 0; r < G->world; ++r) {
        if (r == G->rank) { G->peer_base[r] = G->base; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof h, sizeof h);
        IVOSW_CUDA(cudaIpcOpenMemHandle(&G->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    for (int r = 0; r < G->world; ++r) {
        G->peers.mq[r] = static_cast<double*>(G->peer_base[r]);
        G->peers.flag[r] = reinterpret_cast<unsigned long long*>(static_cast<double*>(G->peer_base[r]) + 2 * (size_t)G->cap);
    }
    G->epoch = G->peers.flag[G->rank] + 2 * 64;
    G->opened = true;
    return IVOSW_OK;
}

int gather_post(ivosw_ctx* c, const double* mq_local_dev, int n_local, int offset, cudaStream_t s) {
    GatherState* G = static_cast<GatherState*>(c->gather_state);
    if (!G || !G->opened) { set_error("gather buffers are not open (ivosw_gather_create / ivosw_gather_open)"); return IVOSW_ERR_STATE; }
    if (offset < 0 || n_local < 0 || offset + n_local > G->cap) { set_error("invalid argument: gather slice outside the buffer"); return IVOSW_ERR_INVALID; }
    gather_post_kernel<<<1, 256, 0, s>>>(G->peers, G->epoch, mq_local_dev, n_local, offset, G->cap, G->world, G->rank);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int gather_wait_pack(ivosw_ctx* c, const double* ann_dev, int T, float* state_dev, double* mq_out_dev, cudaStream_t s) {
    GatherState* G = static_cast<GatherState*>(c->gather_state);
    if (!G || !G->opened) { set_error("gather buffers are not open (ivosw_gather_create / ivosw_gather_open)"); return IVOSW_ERR_STATE; }
    if (T > G->cap) { set_error("invalid argument: T exceeds the gather buffer capacity"); return IVOSW_ERR_INVALID; }
    gather_wait_pack_kernel<<<1, 256, 0, s>>>(G->peers.mq[G->rank], G->peers.flag[G->rank], G->epoch, G->cap, G->world, ann_dev, T,
                                              state_dev, mq_out_dev);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

}  // namespace ivosw
