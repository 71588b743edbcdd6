// Microbenchmark: sustained tcgen05.mma rate of the conv kernel's K block (4 K steps of the x3 split-fp16
// instruction pair:  [D0|D1] += A_hi * [W_hi;W_lo]^T  (N = 256)  and  D1 += A_lo * W_hi^T  (N = 128), M = 128),
// operands resident in shared memory, no TMA, no epilogue: 768 cycles per K block at the nominal rate.
// What is varied is the ISSUE side, because that is what turned out to bound the real kernel:
//   HANDSHAKE 0: back-to-back issue, one commit per K block to a barrier nobody waits on
//   HANDSHAKE 1: the real kernel's ring: wait full[slot] -> 8 MMAs -> commit empty[slot]; a relay thread turns
//                empty[slot] into full[slot] (a producer with zero load latency)
//   ISSUE 0: one thread inside `if (lane == 0)` (operands are per-thread values: ptxas wraps every UTCHMMA/UTCBAR in
//            an ELECT / BRA.U.ANY waterfall loop and moves descriptors with R2UR)
//   ISSUE 1: the whole warp runs the loop with warp-uniform values, elect.sync picks the issuing lane
//   ISSUE 2: single thread, software-pipelined wait: the NEXT block's full barrier is waited on before the last
//            instruction of this block's burst, so no wait sits between two bursts
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu
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
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
    const uint32_t addr = s_u32(b);
    for (uint32_t it = 0;; ++it) {
        uint32_t done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (it > 200000000u) __trap();
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128, STAGE = 2 * A_BYTES + B_BYTES, SLOTS = 3;

template <int ISSUE, int HANDSHAKE, int WAIT_AT>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t done_bar, dummy_bar[4], full_bar[SLOTS], empty_bar[SLOTS];
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < SLOTS * STAGE / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u ^ (uint32_t)(i * 2654435761u & 0x03ff03ffu);   // small fp16 values
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&done_bar)), "r"(1) : "memory");
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&dummy_bar[i])), "r"(1000000) : "memory");
        for (int i = 0; i < SLOTS; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&full_bar[i])), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&empty_bar[i])), "r"(1) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc_256 = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    if (ISSUE < 2 && warp == 0 && (ISSUE == 1 || lane == 0)) {
        int slot = 0; uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            if (HANDSHAKE) {
                mb_wait(&full_bar[slot], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t st = s_u32(smem) + (uint32_t)(slot * STAGE);
            const uint64_t a_hi = sw128_desc(st), a_lo = sw128_desc(st + A_BYTES), b = sw128_desc(st + 2 * A_BYTES);
            const uint32_t d0 = tmem + (uint32_t)((it & 1) * 256);
            if (ISSUE == 0 || elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                    mma(d0, a_hi + adv, b + adv, idesc_256, (it > 1 || k > 0) ? 1u : 0u);
                    mma(d0 + 128, a_lo + adv, b + adv, idesc_128, 1u);
                }
                commit(HANDSHAKE ? &empty_bar[slot] : &dummy_bar[it & 3]);
                if (it == iters - 1) commit(&done_bar);
            }
            if (ISSUE == 1) __syncwarp();
            if (++slot == SLOTS) { slot = 0; ph ^= 1; }
        }
    } else if (ISSUE == 2 && warp == 0 && lane == 0) {
        int slot = 0; uint32_t ph = 0;
        mb_wait(&full_bar[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = s_u32(smem) + (uint32_t)(slot * STAGE);
            const uint64_t a_hi = sw128_desc(st), a_lo = sw128_desc(st + A_BYTES), b = sw128_desc(st + 2 * A_BYTES);
            const uint32_t d0 = tmem + (uint32_t)((it & 1) * 256);
            const int nslot = slot + 1 == SLOTS ? 0 : slot + 1;
            const uint32_t nph = slot + 1 == SLOTS ? ph ^ 1 : ph;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                mma(d0, a_hi + adv, b + adv, idesc_256, (it > 1 || k > 0) ? 1u : 0u);
                if (k == WAIT_AT && it + 1 < iters) {
                    mb_wait(&full_bar[nslot], nph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mma(d0 + 128, a_lo + adv, b + adv, idesc_128, 1u);
            }
            commit(&empty_bar[slot]);
            if (it == iters - 1) commit(&done_bar);
            slot = nslot; ph = nph;
        }
    } else if (HANDSHAKE && warp == 1 && lane == 0) {
        int slot = 0; uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mb_wait(&empty_bar[slot], ph ^ 1);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&full_bar[slot])) : "memory");
            if (++slot == SLOTS) { slot = 0; ph ^= 1; }
        }
    }
    mb_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int ISSUE, int HANDSHAKE, int WAIT_AT = 3>
static void run(int iters, int sms) {
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(rate_kernel<ISSUE, HANDSHAKE, WAIT_AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long* cyc;
    CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        rate_kernel<ISSUE, HANDSHAKE, WAIT_AT><<<sms, 128, smem>>>(iters, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    long long h0;
    CK(cudaMemcpy(&h0, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    const double flops = 4.0 * (2.0 * 128 * 256 * 16 + 2.0 * 128 * 128 * 16) * iters * sms;
    printf("issue %s (wait at K step %d), %s: %.3f ms, %.0f TFLOP/s executed, %.0f cycles per K block\n",
           ISSUE == 2 ? "pipelined wait   " : ISSUE ? "warp + elect.sync" : "single thread   ", WAIT_AT, HANDSHAKE ? "ring handshake" : "free running  ", best,
           flops / best * 1e-9, (double)h0 / iters);
    CK(cudaFree(cyc));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, iters = 20000;
    printf("%s, %d SMs, nominal 768 cycles per K block\n", p.name, sms);
    run<0, 0>(iters, sms);
    run<1, 0>(iters, sms);
    run<0, 1>(iters, sms);
    run<1, 1>(iters, sms);
    run<2, 1, 3>(iters, sms);
    run<2, 1, 2>(iters, sms);
    run<2, 1, 1>(iters, sms);
    run<2, 1, 0>(iters, sms);
    return 0;
}
