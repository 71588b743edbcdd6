// Microbenchmark 2: what sits between the conv kernel's K blocks and the nominal tensor rate?
// Same instruction mix as scripts/umma_rate.cu (per K block: 4 x { [D0|D1] += A_hi [W_hi;W_lo]^T (N = 256),
// D1 += A_lo W_hi^T (N = 128) }, M = 128, both operands in shared memory; 768 cycles at the nominal rate), warp +
// elect.sync issue.  Each MODE isolates one suspect:
//   0  free running, one commit per block to a barrier nobody waits on                     (reference: 768)
//   1  ring handshake: wait full[slot] -> fence -> 8 MMAs -> commit empty[slot]; relay thread = zero-latency producer
//   2  as 1 without tcgen05.fence::after_thread_sync
//   3  free running + a try_wait on an always-complete barrier + fence between bursts       (cost of the wait itself)
//   4  free running, commits go to the ring's empty barriers and the relay runs, issuer never waits   (relay traffic)
//   5  TWO issuing warps, each with its own ring, relay and accumulator pair                (second MMA stream)
//   6  as 1 with two K blocks (16 MMAs) per handshake                                        (bubble per handshake?)
//   7  free running + a copy warp streaming global -> shared with cp.async.bulk (12 x 16 KB in flight) into the
//      operand slots                                                                        (shared-memory port contention)
//   8  as 7 with the ring handshake
//   9  free running + 16 warps polling try_wait on a pending barrier                        (epilogue warps waiting)
//  10  as 5 + the copy warp of 7
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_rate2 umma_rate2.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ bool mb_try(uint64_t* b, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(s_u32(b)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
    for (uint32_t it = 0;; ++it) {
        if (mb_try(b, parity)) return;
        if (it > 100000000u) __trap();
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128, STAGE = 2 * A_BYTES + B_BYTES, SLOTS = 3;
constexpr int CP_CHUNK = 16384, CP_INFLIGHT = 12;
constexpr int THREADS = 640;     // warps 0/2 issue, 1/3 relay (lane 0), 4..19 pollers (mode 9), warp 4 lane 0 copies (7, 8, 10)

struct Bars {
    uint64_t done[2], dummy[4], full[2][SLOTS], empty[2][SLOTS], ready, pending, cp[CP_INFLIGHT];
    uint32_t tmem_slot;
    volatile int stop;
};

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) rate_kernel(int iters, long long* cycles, unsigned long long* copied,
                                                          const uint8_t* gsrc) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ Bars B;
    constexpr bool HANDSHAKE = MODE == 1 || MODE == 2 || MODE == 5 || MODE == 6 || MODE == 8 || MODE == 10;
    constexpr bool FENCE = MODE != 2;
    constexpr bool TWO = MODE == 5 || MODE == 10;
    constexpr bool RELAY = HANDSHAKE || MODE == 4;
    constexpr bool COPY = MODE == 7 || MODE == 8 || MODE == 10;
    constexpr int BURSTS = MODE == 6 ? 2 : 1;
    for (int i = threadIdx.x; i < SLOTS * STAGE / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u ^ (uint32_t)(i * 2654435761u & 0x03ff03ffu);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) mb_init(&B.done[i], 1);
        for (int i = 0; i < 4; ++i) mb_init(&B.dummy[i], 1000000);
        for (int s = 0; s < 2; ++s)
            for (int i = 0; i < SLOTS; ++i) { mb_init(&B.full[s][i], 1); mb_init(&B.empty[s][i], 1); }
        mb_init(&B.ready, 1); mb_init(&B.pending, 1);
        for (int i = 0; i < CP_INFLIGHT; ++i) mb_init(&B.cp[i], 1);
        B.stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mb_arrive(&B.ready);                       // phase 0 of `ready` is complete from now on
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&B.tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    fence_after();
    const uint32_t tmem = B.tmem_slot;
    const uint32_t idesc_256 = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    if (warp == 0 || (TWO && warp == 2)) {
        const int s = warp >> 1;                                   // stream
        int slot = 0; uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            if (HANDSHAKE) { mb_wait(&B.full[s][slot], ph); if (FENCE) fence_after(); }
            if (MODE == 3) { mb_wait(&B.ready, 0); fence_after(); }
            const uint32_t st = s_u32(smem) + (uint32_t)(slot * STAGE);
            const uint64_t a_hi = sw128_desc(st), a_lo = sw128_desc(st + A_BYTES), b = sw128_desc(st + 2 * A_BYTES);
            const uint32_t d0 = tmem + (uint32_t)(TWO ? s * 256 : (it & 1) * 256);
            if (elect_one()) {
#pragma unroll
                for (int r = 0; r < BURSTS; ++r)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        mma(d0, a_hi + adv, b + adv, idesc_256, (it > 1 || k > 0 || r > 0) ? 1u : 0u);
                        mma(d0 + 128, a_lo + adv, b + adv, idesc_128, 1u);
                    }
                commit(RELAY ? &B.empty[s][slot] : &B.dummy[it & 3]);
                if (it == iters - 1) commit(&B.done[s]);
            }
            __syncwarp();
            if (++slot == SLOTS) { slot = 0; ph ^= 1; }
        }
    } else if (RELAY && (warp == 1 || (TWO && warp == 3)) && lane == 0) {
        const int s = warp >> 1;
        int slot = 0; uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mb_wait(&B.empty[s][slot], ph ^ 1);
            mb_arrive(&B.full[s][slot]);
            if (++slot == SLOTS) { slot = 0; ph ^= 1; }
        }
    } else if (COPY && warp == 4 && lane == 0) {
        unsigned long long n = 0;
        const uint8_t* src = gsrc + (size_t)blockIdx.x * CP_INFLIGHT * CP_CHUNK;
        for (uint32_t i = 0; !B.stop; ++i) {
            const int b = i % CP_INFLIGHT;
            if (i >= CP_INFLIGHT) mb_wait(&B.cp[b], ((i / CP_INFLIGHT) - 1) & 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&B.cp[b])), "r"(CP_CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s_u32(smem + b * CP_CHUNK)), "l"(src + b * CP_CHUNK), "r"(CP_CHUNK), "r"(s_u32(&B.cp[b])) : "memory");
            n += CP_CHUNK;
        }
        // drain what is still in flight before the CTA exits
        // (every barrier's last armed phase: wait for it)
        copied[blockIdx.x] = n;
        for (int b = 0; b < CP_INFLIGHT; ++b) {
            // total uses of barrier b so far
            unsigned long long uses = n / CP_CHUNK / CP_INFLIGHT + ((n / CP_CHUNK) % CP_INFLIGHT > (unsigned)b ? 1 : 0);
            if (uses) mb_wait(&B.cp[b], (uint32_t)((uses - 1) & 1));
        }
    } else if (MODE == 9 && warp >= 4) {
        while (!B.stop) { if (mb_try(&B.pending, 0)) break; }
    }
    if (warp == 0 || warp == 2) {                 // (only these wait: pollers / copier must keep running until `stop`)
        mb_wait(&B.done[0], 0);
        if (TWO) mb_wait(&B.done[1], 0);
        fence_after();
        const long long t1 = clock64();
        if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; B.stop = 1; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}


__device__ __forceinline__ bool mb_test(uint64_t* b, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(s_u32(b)), "r"(parity) : "memory");
    return done != 0;
}

// Modes 11 / 12: the ring handshake with a short path from "block landed" to the first MMA of its burst:
//   * the slot loop is unrolled (slot index = compile-time constant), descriptors are base + constant
//   * the last K step of a burst issues the N = 128 instruction first and the N = 256 one last (128 cycles of queued
//     work behind the issuer instead of 64)
//   * EARLY (mode 11): the next block's barrier is TESTED (non-blocking) before the last K step of the current burst;
//     the blocking wait at the top of the next burst only runs if that test failed
template <bool EARLY>
__global__ void __launch_bounds__(THREADS, 1) rate_kernel_opt(int iters, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ Bars B;
    for (int i = threadIdx.x; i < SLOTS * STAGE / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u ^ (uint32_t)(i * 2654435761u & 0x03ff03ffu);
    if (threadIdx.x == 0) {
        mb_init(&B.done[0], 1);
        for (int i = 0; i < SLOTS; ++i) { mb_init(&B.full[0][i], 1); mb_init(&B.empty[0][i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&B.tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    fence_after();
    const uint32_t tmem = B.tmem_slot;
    const uint32_t idesc_256 = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    if (warp == 0) {
        const uint64_t base = sw128_desc(s_u32(smem));
        uint32_t ph = 0; bool ready = false;
        for (int it0 = 0; it0 < iters; it0 += SLOTS) {
#pragma unroll
            for (int slot = 0; slot < SLOTS; ++slot) {
                const int it = it0 + slot;
                if (it < iters) {
                    if (!ready) mb_wait(&B.full[0][slot], ph);
                    fence_after();
                    const uint64_t a_hi = base + (uint64_t)((slot * STAGE) >> 4), a_lo = a_hi + (A_BYTES >> 4), b = a_hi + (2 * A_BYTES >> 4);
                    const uint32_t d0 = tmem + (uint32_t)((it & 1) * 256);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            mma(d0, a_hi + 2 * k, b + 2 * k, idesc_256, (it > 1 || k > 0) ? 1u : 0u);
                            mma(d0 + 128, a_lo + 2 * k, b + 2 * k, idesc_128, 1u);
                        }
                    }
                    __syncwarp();
                    if (EARLY) {
                        const int nslot = (slot + 1) % SLOTS;
                        ready = (it + 1 < iters) && mb_test(&B.full[0][nslot], nslot == 0 ? ph ^ 1 : ph);
                    }
                    if (elect_one()) {
                        mma(d0 + 128, a_lo + 6, b + 6, idesc_128, 1u);
                        mma(d0, a_hi + 6, b + 6, idesc_256, 1u);
                        commit(&B.empty[0][slot]);
                        if (it == iters - 1) commit(&B.done[0]);
                    }
                    __syncwarp();
                }
            }
            ph ^= 1;
        }
    } else if (warp == 1 && lane == 0) {
        int slot = 0; uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mb_wait(&B.empty[0][slot], ph ^ 1);
            mb_arrive(&B.full[0][slot]);
            if (++slot == SLOTS) { slot = 0; ph ^= 1; }
        }
    }
    if (warp == 0) {
        mb_wait(&B.done[0], 0);
        fence_after();
        const long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// Mode 14: the optimised ring of mode 11 fed by REAL loads: a producer thread streams LOAD_BYTES per K block from an
// L2-resident global buffer into the slot (cp.async.bulk, 16 KB pieces, completion on full[slot]) after waiting for
// empty[slot].  64 KB per block is what the conv kernel loads; 48 / 32 KB model designs that deliver fewer bytes per SM.
template <int LOAD_BYTES, bool SPIN, bool NOMMA>
__global__ void __launch_bounds__(THREADS, 1) rate_kernel_fed(int iters, long long* cycles, const uint8_t* gsrc, int region) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ Bars B;
    for (int i = threadIdx.x; i < SLOTS * STAGE / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u ^ (uint32_t)(i * 2654435761u & 0x03ff03ffu);
    if (threadIdx.x == 0) {
        mb_init(&B.done[0], 1);
        for (int i = 0; i < SLOTS; ++i) { mb_init(&B.full[0][i], 1); mb_init(&B.empty[0][i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&B.tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    fence_after();
    const uint32_t tmem = B.tmem_slot;
    const uint32_t idesc_256 = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    if (warp == 0) {
        const uint64_t base = sw128_desc(s_u32(smem));
        uint32_t ph = 0; bool ready = false;
        for (int it0 = 0; it0 < iters; it0 += SLOTS) {
#pragma unroll
            for (int slot = 0; slot < SLOTS; ++slot) {
                const int it = it0 + slot;
                if (it < iters) {
                    if (!ready) { if (SPIN) { while (!mb_test(&B.full[0][slot], ph)) {} } else mb_wait(&B.full[0][slot], ph); }
                    fence_after();
                    const uint64_t a_hi = base + (uint64_t)((slot * STAGE) >> 4), a_lo = a_hi + (A_BYTES >> 4), b = a_hi + (2 * A_BYTES >> 4);
                    const uint32_t d0 = tmem + (uint32_t)((it & 1) * 256);
                    if (!NOMMA && elect_one()) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            mma(d0, a_hi + 2 * k, b + 2 * k, idesc_256, (it > 1 || k > 0) ? 1u : 0u);
                            mma(d0 + 128, a_lo + 2 * k, b + 2 * k, idesc_128, 1u);
                        }
                    }
                    __syncwarp();
                    const int nslot = (slot + 1) % SLOTS;
                    ready = (it + 1 < iters) && mb_test(&B.full[0][nslot], nslot == 0 ? ph ^ 1 : ph);
                    if (elect_one()) {
                        if (!NOMMA) {
                            mma(d0 + 128, a_lo + 6, b + 6, idesc_128, 1u);
                            mma(d0, a_hi + 6, b + 6, idesc_256, 1u);
                        }
                        commit(&B.empty[0][slot]);
                        if (it == iters - 1) commit(&B.done[0]);
                    }
                    __syncwarp();
                }
            }
            ph ^= 1;
        }
    } else if (warp == 1 && lane == 0) {
        int slot = 0; uint32_t ph = 0;
        const uint8_t* src = gsrc + (size_t)blockIdx.x * region;
        uint32_t off = 0;
        for (int it = 0; it < iters; ++it) {
            if (SPIN) { while (!mb_test(&B.empty[0][slot], ph ^ 1)) {} } else mb_wait(&B.empty[0][slot], ph ^ 1);
            if (LOAD_BYTES == 0) mb_arrive(&B.full[0][slot]);
            else asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&B.full[0][slot])), "r"(LOAD_BYTES) : "memory");
#pragma unroll
            for (int c = 0; c < LOAD_BYTES / CP_CHUNK; ++c) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s_u32(smem + slot * STAGE + c * CP_CHUNK)), "l"(src + off), "r"(CP_CHUNK), "r"(s_u32(&B.full[0][slot])) : "memory");
                off += CP_CHUNK; if (off >= (uint32_t)region) off = 0;
            }
            if (++slot == SLOTS) { slot = 0; ph ^= 1; }
        }
    }
    if (warp == 0) {
        mb_wait(&B.done[0], 0);
        fence_after();
        const long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int MODE>
static void run(const char* what, int iters, int sms, const uint8_t* gsrc) {
    const int smem = SLOTS * STAGE + 2048;
    CK(cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long* cyc; unsigned long long* cp;
    CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    CK(cudaMalloc(&cp, sizeof(unsigned long long) * sms));
    CK(cudaMemset(cp, 0, sizeof(unsigned long long) * sms));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        rate_kernel<MODE><<<sms, THREADS, smem>>>(iters, cyc, cp, gsrc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    long long h0; unsigned long long c0;
    CK(cudaMemcpy(&h0, cyc, sizeof h0, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&c0, cp, sizeof c0, cudaMemcpyDeviceToHost));
    const int streams = (MODE == 5 || MODE == 10) ? 2 : 1;
    const int bursts = MODE == 6 ? 2 : 1;
    const double blocks = (double)iters * streams * bursts;
    const double flops = 4.0 * (2.0 * 128 * 256 * 16 + 2.0 * 128 * 128 * 16) * blocks * sms;
    printf("mode %2d %-58s %8.3f ms %6.0f TFLOP/s %7.1f cycles/K-block  copy %.1f B/clk\n", MODE, what, best,
           flops / best * 1e-9, (double)h0 / blocks, (double)c0 / (double)h0);
    CK(cudaFree(cyc)); CK(cudaFree(cp));
}

template <bool EARLY>
static void run_opt(const char* what, int iters, int sms) {
    const int smem = SLOTS * STAGE + 2048;
    CK(cudaFuncSetAttribute(rate_kernel_opt<EARLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long* cyc;
    CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        rate_kernel_opt<EARLY><<<sms, THREADS, smem>>>(iters, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    long long h0;
    CK(cudaMemcpy(&h0, cyc, sizeof h0, cudaMemcpyDeviceToHost));
    const double flops = 4.0 * (2.0 * 128 * 256 * 16 + 2.0 * 128 * 128 * 16) * iters * sms;
    printf("mode %2d %-58s %8.3f ms %6.0f TFLOP/s %7.1f cycles/K-block\n", EARLY ? 11 : 12, what, best, flops / best * 1e-9,
           (double)h0 / iters);
    CK(cudaFree(cyc));
}

template <int LOAD_BYTES, bool SPIN = false, bool NOMMA = false>
static void run_fed(int iters, int sms, const uint8_t* gsrc, int region) {
    const int smem = SLOTS * STAGE + 2048;
    CK(cudaFuncSetAttribute(rate_kernel_fed<LOAD_BYTES, SPIN, NOMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long* cyc;
    CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        rate_kernel_fed<LOAD_BYTES, SPIN, NOMMA><<<sms, THREADS, smem>>>(iters, cyc, gsrc, region);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    long long h0;
    CK(cudaMemcpy(&h0, cyc, sizeof h0, cudaMemcpyDeviceToHost));
    const double flops = 4.0 * (2.0 * 128 * 256 * 16 + 2.0 * 128 * 128 * 16) * iters * sms;
    printf("mode 14%s%s optimised ring fed by bulk loads, %2d KB per K block, %4d KB region per CTA %8.3f ms %6.0f TFLOP/s %7.1f cycles/K-block  load %.1f B/clk/SM\n",
           SPIN ? " SPIN" : "", NOMMA ? " NO-MMA" : "", LOAD_BYTES / 1024, region / 1024, best, flops / best * 1e-9, (double)h0 / iters, (double)LOAD_BYTES * iters / (double)h0);
    CK(cudaFree(cyc));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, iters = 20000;
    printf("%s, %d SMs, nominal 768 cycles per K block\n", p.name, sms);
    uint8_t* gsrc;
    const size_t gbytes = (size_t)sms * CP_INFLIGHT * CP_CHUNK;
    CK(cudaMalloc(&gsrc, gbytes));
    CK(cudaMemset(gsrc, 0x2c, gbytes));
    run<0>("free running", iters, sms, gsrc);
    run<1>("ring handshake", iters, sms, gsrc);
    run<2>("ring handshake, no fence::after", iters, sms, gsrc);
    run<3>("free running + try_wait(complete) + fence per burst", iters, sms, gsrc);
    run<4>("free running, commits to ring + relay, no wait", iters, sms, gsrc);
    run<5>("two issuing warps, ring handshake each", iters, sms, gsrc);
    run<6>("ring handshake, 16 MMAs per handshake", iters, sms, gsrc);
    run<7>("free running + bulk copies into the slots", iters, sms, gsrc);
    run<8>("ring handshake + bulk copies into the slots", iters, sms, gsrc);
    run<9>("free running + 16 warps polling a pending barrier", iters, sms, gsrc);
    run<10>("two issuing warps + bulk copies", iters, sms, gsrc);
    run_opt<true>("ring handshake, unrolled slots, N=256 last, early test", iters, sms);
    run_opt<false>("ring handshake, unrolled slots, N=256 last", iters, sms);
    {
        uint8_t* big;
        const int region = 512 * 1024;                      // 148 x 512 KB = 74 MB: L2-resident
        CK(cudaMalloc(&big, (size_t)sms * region));
        CK(cudaMemset(big, 0x2c, (size_t)sms * region));
        run_fed<65536>(iters, sms, big, region);
        run_fed<49152>(iters, sms, big, region);
        run_fed<32768>(iters, sms, big, region);
        run_fed<16384>(iters, sms, big, region);
        run_fed<65536>(iters, sms, big, 192 * 1024);
        run_fed<65536, true>(iters, sms, big, region);
        run_fed<49152, true>(iters, sms, big, region);
        run_fed<0, false, true>(iters, sms, big, region);
        run_fed<0, true, true>(iters, sms, big, region);
        run_fed<65536, false, true>(iters, sms, big, region);
        run_fed<65536, true, true>(iters, sms, big, region);
    }
    return 0;
}
