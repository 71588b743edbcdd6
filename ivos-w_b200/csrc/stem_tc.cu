// AssessNet stem on the tensor cores, with the max-pool fused (models/assessment.py:54-57):
//     x = conv1(f) + conv1_p(p)  ->  bn1  ->  relu  ->  maxpool(3, 2, 1)
// evaluated as ONE 4-input-channel 7x7 stride-2 convolution over the NHWC ROI crop
// (channels 0..2 normalised RGB, 3 probability).
//
// GEMM view per CTA tile: M = 128 = one full output row of c1 (ow = 0..127), N = 64 channels,
// K = 7 filter rows x 32 (7 kw x 4 channels + 4 zero pad) = 224.
//
// The A operand is never built.  The ROI sampler wrote the crop as split-fp16 planes on a zero-bordered
// canvas (roi.cu: 8 bytes per pixel per plane, image at offset (3, 3)), so in a raw canvas row the K slab
// "filter taps kw = 2*sp, 2*sp + 1 of output column ow" is the 16 bytes at offset 16 * (ow + sp).  That IS
// the canonical no-swizzle K-major UMMA layout with overlapping core matrices:
//     8 consecutive output columns = 128 contiguous bytes (rows 16 bytes apart),
//     next 8 columns: +128 bytes (SBO), next K slab: +16 bytes (LBO).
// One TMA bulk copy per canvas row pair lands in a shared-memory ring, and tcgen05.mma reads the 7 filter
// rows of an output row straight from it (descriptor start = row address + 32 * k16-step).  A canvas row
// pair is loaded once and used by four consecutive output rows.
// Weights (hi/lo, 56 KB) are packed on the host into the matching layout — per K slab the 8 row groups of
// W_hi followed by the 8 of W_lo, so that A_hi * [W_hi ; W_lo]^T is one N = 128 instruction — and stay
// resident in shared memory for the whole kernel.
//
// A CTA walks consecutive c1 rows of one image band, so the epilogue can fuse the 3x3/2 max-pool:
// each BN+ReLU'd c1 row goes to a shared-memory row buffer, is pooled horizontally, and a running
// vertical max kept in registers emits one pooled row every second c1 row, already in the split-fp16
// form the bottleneck convolutions consume.  c1 (537 MB per 128 units in fp32) never touches HBM.
//
// Algorithmic work 0.411 GFLOP per unit; x3 MMA terms, K padded 196 -> 224.
#include "ivosw_internal.h"

namespace ivosw {

namespace stemtc {

constexpr int THREADS = 384;          // warp 0: MMA issuer + TMEM owner; warp 1: TMA producer; warps 4-11: epilogue
constexpr int KPAD = 224;             // 7 * 32
constexpr int SLABS = KPAD / 8;       // 28 K slabs of 8 elements (16 bytes)
constexpr int B_SLAB = 16 * 128;      // per slab: 8 row groups of W_hi then 8 of W_lo, 128 B each      : 2048
constexpr int B_BYTES = SLABS * B_SLAB;                  // 57344
constexpr int ROW_BYTES = CROP_PW * 8;                   // one canvas row of one plane                   : 2112
constexpr int PAIR_PLANE = 2 * ROW_BYTES;                // canvas rows 2p, 2p + 1 of one plane           : 4224 (= 33 * 128)
constexpr int PAIR_BYTES = 2 * PAIR_PLANE;               // hi plane then lo plane                        : 8448
constexpr int RING = 8;                                  // canvas row pairs resident (an output row reads 4)
constexpr int ROWBUF = 128 * 64 * 4;                     // one c1 row, fp32, [ch/4][ow] float4           : 32768
constexpr int OFF_B = 0;
constexpr int OFF_RING = OFF_B + B_BYTES;                // 57344
constexpr int OFF_ROW = OFF_RING + RING * PAIR_BYTES;    // 124928
constexpr int OFF_BAR = OFF_ROW + ROWBUF;                // 157696
constexpr int SMEM_TOTAL = OFF_BAR + 256 + 512 + 1024;   // barriers, scale/shift, alignment slack
constexpr long long WAIT_TIMEOUT = 4000000000ll;
static_assert(CROP_PH >= 2 * 127 + 8 && (CROP_PH % 2) == 0, "canvas must hold row pair oh + 3 of the last output row");
static_assert(PAIR_PLANE % 128 == 0, "row pair planes keep the ring 128-byte aligned");

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
    const uint32_t addr = s_u32(b);
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        uint32_t done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if ((it & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > WAIT_TIMEOUT) __trap();
        }
    }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// no-swizzle K-major descriptor: core matrices of 8 rows x 16 bytes stored contiguously (128 B);
// lbo = byte distance between core matrices adjacent in K, sbo = between 8-row groups (M / N).
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;        // descriptor version (sm_100); layout type 0 = SWIZZLE_NONE
    return d;
}

struct Params {
    const uint8_t* crop_hi;    // [B][CROP_PH][CROP_PW] pixels of 4 fp16 channels (8 bytes)
    const uint8_t* crop_lo;
    const uint4* wpack;        // B_BYTES: the shared-memory image of the weights
    const float* scale;
    const float* shift;
    __half* out_hi;            // [B][64][64][64] NHWC
    __half* out_lo;
    int n_items;               // B * bands
    int bands;                 // row bands per image
    int terms;                 // 3 or 1
    unsigned long long* sat_count;   // fp16 range guard events (ivosw_conv_saturation_count)
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(THREADS, 1) stem_tc_kernel(const __grid_constant__ Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* p_full = bars;                  // [RING] canvas row pair landed            (TMA bytes)
    uint64_t* p_empty = bars + RING;          // [RING] its last reader has finished      (tcgen05.commit)
    uint64_t* t_full = bars + 2 * RING;       // [2] MMA -> epilogue
    uint64_t* t_empty = bars + 2 * RING + 2;  // [2] epilogue -> MMA                      (count 8)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * RING + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool x3 = P.terms == 3;

    // weights: plain 16-byte copies of the pre-packed smem image
    for (int i = threadIdx.x; i < B_BYTES / 16; i += THREADS)
        reinterpret_cast<uint4*>(smem + OFF_B)[i] = __ldg(P.wpack + i);
    float* sss = reinterpret_cast<float*>(smem + OFF_BAR + 256);        // [64 scale][64 shift] (bn1 folded)
    if (threadIdx.x < 128) sss[threadIdx.x] = threadIdx.x < 64 ? P.scale[threadIdx.x] : P.shift[threadIdx.x - 64];
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) { mb_init(&p_full[i], 1); mb_init(&p_empty[i], 1); }
        mb_init(&t_full[0], 1); mb_init(&t_full[1], 1); mb_init(&t_empty[0], 8); mb_init(&t_empty[1], 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();              // weight image written with generic stores, read by the tensor core
    tc_before();
    __syncthreads();
    tc_after();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");              // (programmatic dependent launch, see conv_tc.cu)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int rows_per_band = 64 / P.bands;          // pooled rows per band

    if (warp == 1) {
        // ------------------------------------------------ TMA producer: canvas row pairs, in the order they are first read
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            const uint32_t tx = x3 ? PAIR_BYTES : PAIR_PLANE;
            for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
                const int img = item / P.bands, band = item % P.bands;
                const int r_first = band == 0 ? 0 : 2 * band * rows_per_band - 1;      // first c1 row of the band
                const int n_rows = 2 * rows_per_band + (band == 0 ? 0 : 1);
                const size_t plane0 = (size_t)img * CROP_PH * ROW_BYTES;
                for (int pr = r_first; pr < r_first + n_rows + 3; ++pr) {           // output row oh reads pairs oh .. oh + 3
                    mb_wait(&p_empty[slot], ph ^ 1);
                    uint8_t* dst = smem + OFF_RING + slot * PAIR_BYTES;
                    mb_expect_tx(&p_full[slot], tx);
                    bulk_load(dst, P.crop_hi + plane0 + (size_t)pr * PAIR_PLANE, PAIR_PLANE, &p_full[slot]);
                    if (x3) bulk_load(dst + PAIR_PLANE, P.crop_lo + plane0 + (size_t)pr * PAIR_PLANE, PAIR_PLANE, &p_full[slot]);
                    if (++slot == RING) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 0) {
        // ------------------------------------------------ MMA issuer (whole warp, elect.sync; see conv_tc.cu)
        const uint32_t idesc_n128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_n64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t ring = s_u32(smem + OFF_RING), bw = s_u32(smem + OFF_B);
        int slot = 0; uint32_t ph = 0;            // ring position of pair `oh` (the oldest pair of the current output row)
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
            const int n_rows = 2 * rows_per_band + ((item % P.bands) == 0 ? 0 : 1);   // c1 rows of this band
            // pairs oh .. oh + 2 of the band's first row; every row then waits for one more pair
            {
                int s2 = slot; uint32_t p2 = ph;
                for (int i = 0; i < 3; ++i) { mb_wait(&p_full[s2], p2); if (++s2 == RING) { s2 = 0; p2 ^= 1; } }
            }
            for (int r = 0; r < n_rows; ++r) {
                {
                    int s3 = slot + 3; uint32_t p3 = ph;
                    if (s3 >= RING) { s3 -= RING; p3 ^= 1; }
                    mb_wait(&p_full[s3], p3);
                }
                mb_wait(&t_empty[acc], acc_phase ^ 1);
                tc_after();
                const uint32_t d0 = tmem_base + (uint32_t)(acc * 128), d1 = d0 + 64;
                if (elect_one()) {
#pragma unroll
                    for (int kh = 0; kh < 7; ++kh) {
                        int sk = slot + (kh >> 1);
                        if (sk >= RING) sk -= RING;
                        const uint32_t row_hi = ring + (uint32_t)(sk * PAIR_BYTES + (kh & 1) * ROW_BYTES);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            // K = 16 step: slabs 2j, 2j + 1 of this filter row; A straight from the canvas row
                            const uint64_t da_hi = make_desc(row_hi + j * 32, 16, 128);
                            const uint64_t da_lo = make_desc(row_hi + PAIR_PLANE + j * 32, 16, 128);
                            const uint64_t db = make_desc(bw + (uint32_t)((kh * 2 + j) * 2 * B_SLAB), B_SLAB, 128);
                            const uint32_t accum = (kh > 0 || j > 0) ? 1u : 0u;
                            if (x3) {
                                mma_f16(d0, da_hi, db, idesc_n128, accum);      // [D0 | D1] += A_hi * [W_hi ; W_lo]^T
                                mma_f16(d1, da_lo, db, idesc_n64, 1u);          //       D1  += A_lo * W_hi^T
                            } else {
                                mma_f16(d0, da_hi, db, idesc_n64, accum);
                            }
                        }
                    }
                    mma_commit(&p_empty[slot]);                   // pair `oh` has no reader after this row
                    if (r == n_rows - 1) {                        // band end: the three younger pairs are done too
                        int s2 = slot;
                        for (int i = 0; i < 3; ++i) { if (++s2 == RING) s2 = 0; mma_commit(&p_empty[s2]); }
                    }
                    mma_commit(&t_full[acc]);
                }
                __syncwarp();
                if (++slot == RING) { slot = 0; ph ^= 1; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            for (int i = 0; i < 3; ++i) { if (++slot == RING) { slot = 0; ph ^= 1; } }     // skip the band's trailing pairs
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue + fused max-pool (8 warps)
        // phase 1: warp = (TMEM lane quarter, channel half): BN + ReLU of 32 channels of one c1 column -> row buffer
        // phase 2: thread = (pooled column q, 16-channel quarter): horizontal 3-max, running vertical max, store
        const int quarter = warp & 3;
        const int ow = quarter * 32 + lane;            // TMEM lane = c1 column
        const int e = warp - 4;                        // 0..7
        const int cc = (e >> 2) * 32;                  // channel half of phase 1
        const int et = e * 32 + lane;                  // 0..255
        const int q = et & 63, cq = et >> 6;           // pooled column, channel quarter of phase 2
        float4* rowbuf = reinterpret_cast<float4*>(smem + OFF_ROW);   // [16 channel quads][128 ow]
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
            const int img = item / P.bands, band = item % P.bands;
            const int p0 = band * rows_per_band;
            const int r_first = band == 0 ? 0 : 2 * p0 - 1;
            const int n_rows = 2 * rows_per_band + (band == 0 ? 0 : 1);
            // running vertical max: after an odd c1 row it holds that row (first row of the next pooled
            // row), after an even row the max of rows (2p-1, 2p)
            float state[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) state[k] = -INFINITY;
            for (int r = 0; r < n_rows; ++r) {
                const int oh = r_first + r;
                mb_wait(&t_full[acc], acc_phase);
                tc_after();
                const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 128);
                {
                    uint32_t r0[32], r1[32];
                    ld32(t0 + cc, r0);
                    if (x3) ld32(t0 + 64 + cc, r1);
                    ld_wait();
                    tc_before();
                    __syncwarp();
                    if (lane == 0) mb_arrive(&t_empty[acc]);
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        float o[4];
                        const float4 sc4 = *reinterpret_cast<const float4*>(sss + cc + k4 * 4);
                        const float4 sh4 = *reinterpret_cast<const float4*>(sss + 64 + cc + k4 * 4);
                        const float scv[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, shv[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float a = __uint_as_float(r0[k4 * 4 + u]);
                            if (x3) a = fmaf(__uint_as_float(r1[k4 * 4 + u]), 1.0f / 2048.0f, a);
                            o[u] = fmaxf(fmaf(a, scv[u], shv[u]), 0.f);    // bn1 + relu
                        }
                        rowbuf[((cc >> 2) + k4) * 128 + ow] = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");       // c1 row complete in rowbuf
                // horizontal 3-max for pooled column q, 16 channels
                float hp[16];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4* rb = rowbuf + (cq * 4 + k4) * 128;
                    float4 m = rb[2 * q];
                    const float4 b = rb[2 * q + 1];
                    m.x = fmaxf(m.x, b.x); m.y = fmaxf(m.y, b.y); m.z = fmaxf(m.z, b.z); m.w = fmaxf(m.w, b.w);
                    if (q > 0) {
                        const float4 a = rb[2 * q - 1];
                        m.x = fmaxf(m.x, a.x); m.y = fmaxf(m.y, a.y); m.z = fmaxf(m.z, a.z); m.w = fmaxf(m.w, a.w);
                    }
                    hp[k4 * 4] = m.x; hp[k4 * 4 + 1] = m.y; hp[k4 * 4 + 2] = m.z; hp[k4 * 4 + 3] = m.w;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");       // rowbuf free for the next c1 row
                if ((oh & 1) == 0) {                                  // row 2p: state = max(row 2p-1, row 2p)
#pragma unroll
                    for (int k = 0; k < 16; ++k) state[k] = fmaxf(state[k], hp[k]);
                } else {                                              // row 2p+1 closes pooled row p
                    const int p = oh >> 1;
                    if (p >= p0) {                                    // (the band's extra leading row only primes carry)
                        uint32_t oh_[8], ol_[8];
#pragma unroll
                        for (int k = 0; k < 16; k += 2) {
                            const float a = fminf(fmaxf(state[k], hp[k]), 65504.f), b = fminf(fmaxf(state[k + 1], hp[k + 1]), 65504.f);
                            const __half2 h = __floats2half2_rn(a, b);
                            const float2 hf = __half22float2(h);
                            oh_[k >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                            if (((oh_[k >> 1] & 0x7FFF7FFFu) + 0x04010401u) & 0x80008000u) atomicAdd(P.sat_count, 1ull);   // clamped to 65504
                            ol_[k >> 1] = pack2((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
                        }
                        const size_t off = (((size_t)img * 64 + p) * 64 + q) * 64 + cq * 16;
                        uint4* dh = reinterpret_cast<uint4*>(P.out_hi + off);
                        uint4* dl = reinterpret_cast<uint4*>(P.out_lo + off);
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            dh[u] = make_uint4(oh_[4 * u], oh_[4 * u + 1], oh_[4 * u + 2], oh_[4 * u + 3]);
                            dl[u] = make_uint4(ol_[4 * u], ol_[4 * u + 1], ol_[4 * u + 2], ol_[4 * u + 3]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k) state[k] = hp[k];
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tc_before();
    __syncthreads();
    if (warp == 0) {
        tc_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

}  // namespace stemtc

// host: pack the [64][7][7][4] stem weight into the smem image: per K slab 8 row groups of W_hi, then 8 of W_lo
int stem_tc_pack(ivosw_ctx* c, const float* w_ohwi /*[64][196]*/) {
    using namespace stemtc;
    std::vector<__half> img((size_t)B_BYTES / 2, __float2half_rn(0.f));
    for (int n = 0; n < 64; ++n)
        for (int kh = 0; kh < 7; ++kh)
            for (int kw = 0; kw < 7; ++kw)
                for (int ch = 0; ch < 4; ++ch) {
                    const float w = w_ohwi[(size_t)n * 196 + (kh * 7 + kw) * 4 + ch];
                    const int k = kh * 32 + kw * 4 + ch;
                    const size_t slab = (size_t)(k >> 3) * (B_SLAB / 2);                       // in halves
                    const size_t in_group = (size_t)(n & 7) * 8 + (k & 7);
                    const __half h = __float2half_rn(w);
                    img[slab + (size_t)(n >> 3) * 64 + in_group] = h;
                    img[slab + (size_t)(8 + (n >> 3)) * 64 + in_group] = __float2half_rn((w - __half2float(h)) * 2048.0f);
                }
    if (!c->stem_wpack) IVOSW_CUDA(cudaMalloc(&c->stem_wpack, (size_t)B_BYTES));
    IVOSW_CUDA(cudaMemcpy(c->stem_wpack, img.data(), (size_t)B_BYTES, cudaMemcpyHostToDevice));
    IVOSW_CUDA(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    return IVOSW_OK;
}

int launch_stem_tc(ivosw_ctx* c, int B, const SplitAct& out, int terms, cudaStream_t s) {
    using namespace stemtc;
    // row bands per image: minimise waves * (c1 rows per band)
    int best = 1; long long best_cost = -1;
    for (int nb = 1; nb <= 16; nb *= 2) {
        const long long items = (long long)B * nb, waves = (items + c->sm_count - 1) / c->sm_count;
        const long long cost = waves * (128 / nb + 1);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = nb; }
    }
    Params P;
    P.crop_hi = (const uint8_t*)c->crop_hi.p; P.crop_lo = (const uint8_t*)c->crop_lo.p; P.wpack = (const uint4*)c->stem_wpack;
    P.scale = c->stem_scale; P.shift = c->stem_shift;
    P.out_hi = out.hi; P.out_lo = out.lo;
    P.bands = best; P.n_items = B * best; P.terms = terms;
    P.sat_count = c->sat_count;
    const int grid = P.n_items < c->sm_count ? P.n_items : c->sm_count;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_TOTAL; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = !(getenv("IVOSW_PDL") && atoi(getenv("IVOSW_PDL")) == 0);
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    IVOSW_CUDA(cudaLaunchKernelEx(&cfg, stem_tc_kernel, P));
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

}  // namespace ivosw
