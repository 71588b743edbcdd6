// AssessNet stem on the tensor cores, with the max-pool fused (models/assessment.py:54-57):
//     x = conv1(f) + conv1_p(p)  ->  bn1  ->  relu  ->  maxpool(3, 2, 1)
// evaluated as ONE 4-input-channel 7x7 stride-2 convolution over the NHWC ROI crop
// (channels 0..2 normalised RGB, 3 probability).
//
// GEMM view per CTA tile: M = 128 = one full output row of c1 (ow = 0..127), N = 64 channels,
// K = 7 filter rows x 32 (7 kw x 4 channels + 4 zero pad) = 224.  The 4-channel pixels are far too
// narrow for a TMA box, so the A operand is built by the SM: the ROI sampler already wrote the crop as
// split-fp16 planes on a zero-bordered canvas (roi.cu), so a K slab (two adjacent pixels x 4 channels)
// is ONE aligned 16-byte load per plane; "builder" warps copy slabs into the canonical no-swizzle
// K-major UMMA layout (8 x 16 B core matrices) and a proxy fence hands the tile to tcgen05.mma.
// Weights (hi/lo, 56 KB) are packed on the host into the same layout and stay resident in shared
// memory for the whole kernel.
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

constexpr int THREADS = 512;          // warp 0: MMA issuer + TMEM owner; warps 4-7: epilogue; warps 8-15: builders
constexpr int KPAD = 224;             // 7 * 32
constexpr int SLABS = KPAD / 8;       // 28 K slabs of 8 elements (16 bytes)
constexpr int A_PLANE = SLABS * 16 * 128;   // 128 rows -> 16 row groups x 128 B core matrices : 57344
constexpr int B_PLANE = SLABS * 8 * 128;    // 64 rows  ->  8 row groups                       : 28672
constexpr int ROWBUF = 128 * 64 * 4;        // one c1 row, fp32, [ch/4][ow] float4              : 32768
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + 2 * A_PLANE;               // 114688
constexpr int OFF_ROW = OFF_B + 2 * B_PLANE;             // 172032
constexpr int OFF_BAR = OFF_ROW + ROWBUF;                // 204800
constexpr int SMEM_TOTAL = OFF_BAR + 128 + 512 + 1024;     // barriers, scale/shift, alignment slack
constexpr long long WAIT_TIMEOUT = 4000000000ll;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory");
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
    const uint4* crop_hi;      // [B][CROP_PH][CROP_PW / 2] : two pixels x 4 channels fp16 per 16 bytes
    const uint4* crop_lo;
    const uint4* wpack;        // 2 * B_PLANE bytes: hi plane then lo plane, already in the smem image layout
    const float* scale;
    const float* shift;
    __half* out_hi;            // [B][64][64][64] NHWC
    __half* out_lo;
    int n_items;               // B * bands
    int bands;                 // row bands per image
    int terms;                 // 3 or 1
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(THREADS, 1) stem_tc_kernel(const __grid_constant__ Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    // The A tile is built and consumed in two K parts (filter rows 0-3 = K steps 0-7, rows 4-6 = steps 8-13),
    // each with its own full/empty pair, so building part 0 of the next c1 row overlaps the MMAs on part 1.
    uint64_t* a_full = bars;          // [2] builders -> MMA      (count 4: one arrive per builder warp of the part)
    uint64_t* a_empty = bars + 2;     // [2] MMA -> builders      (tcgen05.commit)
    uint64_t* t_full = bars + 4;      // [2] MMA -> epilogue
    uint64_t* t_empty = bars + 6;     // [2] epilogue -> MMA      (count 4)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool x3 = P.terms == 3;

    // weights: plain 16-byte copies of the pre-packed smem image
    for (int i = threadIdx.x; i < 2 * B_PLANE / 16; i += THREADS)
        reinterpret_cast<uint4*>(smem + OFF_B)[i] = __ldg(P.wpack + i);
    float* sss = reinterpret_cast<float*>(smem + OFF_BAR + 128);        // [64 scale][64 shift] (bn1 folded)
    if (threadIdx.x < 128) sss[threadIdx.x] = threadIdx.x < 64 ? P.scale[threadIdx.x] : P.shift[threadIdx.x - 64];
    if (threadIdx.x == 0) {
        mb_init(&a_full[0], 4); mb_init(&a_full[1], 4); mb_init(&a_empty[0], 1); mb_init(&a_empty[1], 1);
        mb_init(&t_full[0], 1); mb_init(&t_full[1], 1); mb_init(&t_empty[0], 4); mb_init(&t_empty[1], 4);
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

    if (warp == 0) {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a_hi = s_u32(smem + OFF_A), a_lo = a_hi + A_PLANE;
            const uint32_t b_hi = s_u32(smem + OFF_B), b_lo = b_hi + B_PLANE;
            uint32_t a_phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
                const int n_rows = 2 * rows_per_band + ((item % P.bands) == 0 ? 0 : 1);   // c1 rows of this band
                for (int r = 0; r < n_rows; ++r) {
                    mb_wait(&t_empty[acc], acc_phase ^ 1);
                    const uint32_t d0 = tmem_base + (uint32_t)(acc * 128), d1 = d0 + 64;
#pragma unroll 1
                    for (int part = 0; part < 2; ++part) {
                        mb_wait(&a_full[part], a_phase);
                        tc_after();
                        const int ks0 = part == 0 ? 0 : 8, ks1 = part == 0 ? 8 : KPAD / 16;
#pragma unroll 1
                        for (int ks = ks0; ks < ks1; ++ks) {
                            // one K=16 step = two 8-element slabs: A slab stride 2048 B, B slab stride 1024 B
                            const uint64_t da_hi = make_desc(a_hi + ks * 4096, 2048, 128), da_lo = make_desc(a_lo + ks * 4096, 2048, 128);
                            const uint64_t db_hi = make_desc(b_hi + ks * 2048, 1024, 128), db_lo = make_desc(b_lo + ks * 2048, 1024, 128);
                            mma_f16(d0, da_hi, db_hi, idesc, ks > 0 ? 1u : 0u);
                            if (x3) {
                                mma_f16(d1, da_hi, db_lo, idesc, ks > 0 ? 1u : 0u);
                                mma_f16(d1, da_lo, db_hi, idesc, 1u);
                            }
                        }
                        mma_commit(&a_empty[part]);
                    }
                    mma_commit(&t_full[acc]);
                    a_phase ^= 1;
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 8) {
        // ------------------------------------------------ A builders (256 threads)
        // thread -> one c1 column (ow) and one K part: part 0 = filter rows 0..3, part 1 = rows 4..6.
        // Filter row kh of output (oh, ow) is canvas row 2*oh + kh, canvas pixels 2*ow .. 2*ow + 7: four
        // aligned 16-byte slabs per plane.  The first two filter rows are loaded before waiting for the slot.
        const int bt = threadIdx.x - 256;
        const int ow = bt & 127, part = bt >> 7;
        const int kh0 = part * 4, nkh = part == 0 ? 4 : 3;
        uint32_t e_phase = 0;
        const uint32_t sA = (uint32_t)((ow >> 3) * 128 + (ow & 7) * 16);
        auto load_row = [&](const uint4* ph, const uint4* pl, int oh, int kh, uint4* vh, uint4* vl) {
            const size_t idx = ((size_t)(2 * oh + kh) * CROP_PW + 2 * ow) >> 1;
#pragma unroll
            for (int sp = 0; sp < 4; ++sp) {
                vh[sp] = __ldg(ph + idx + sp);
                if (x3) vl[sp] = __ldg(pl + idx + sp);
            }
        };
        auto store_row = [&](int kh, const uint4* vh, const uint4* vl) {
#pragma unroll
            for (int sp = 0; sp < 4; ++sp) {
                const uint32_t off = (uint32_t)((kh * 4 + sp) * 2048) + sA;
                *reinterpret_cast<uint4*>(smem + OFF_A + off) = vh[sp];
                if (x3) *reinterpret_cast<uint4*>(smem + OFF_A + A_PLANE + off) = vl[sp];
            }
        };
        for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
            const int img = item / P.bands, band = item % P.bands;
            const int p0 = band * rows_per_band;
            const int r_first = band == 0 ? 0 : 2 * p0 - 1;              // first c1 row of the band
            const int n_rows = 2 * rows_per_band + (band == 0 ? 0 : 1);
            const uint4* ph = P.crop_hi + (size_t)img * CROP_PH * CROP_PW / 2;
            const uint4* pl = P.crop_lo + (size_t)img * CROP_PH * CROP_PW / 2;
            for (int r = 0; r < n_rows; ++r) {
                const int oh = r_first + r;
                uint4 ah[4], al[4], bh[4], bl[4];
                load_row(ph, pl, oh, kh0, ah, al);
                load_row(ph, pl, oh, kh0 + 1, bh, bl);
                mb_wait(&a_empty[part], e_phase ^ 1);
                store_row(kh0, ah, al);
                store_row(kh0 + 1, bh, bl);
                load_row(ph, pl, oh, kh0 + 2, ah, al);
                if (nkh == 4) load_row(ph, pl, oh, kh0 + 3, bh, bl);
                store_row(kh0 + 2, ah, al);
                if (nkh == 4) store_row(kh0 + 3, bh, bl);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mb_arrive(&a_full[part]);
                e_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue + fused max-pool (128 threads)
        const int quarter = warp & 3;
        const int ow = quarter * 32 + lane;            // TMEM lane = c1 column
        const int et = (warp - 4) * 32 + lane;         // 0..127
        const int q = et >> 1, chh = (et & 1) * 32;    // pooled column, channel half handled in the pooling phase
        float4* rowbuf = reinterpret_cast<float4*>(smem + OFF_ROW);   // [16 channel quads][128 ow]
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
            const int img = item / P.bands, band = item % P.bands;
            const int p0 = band * rows_per_band;
            const int r_first = band == 0 ? 0 : 2 * p0 - 1;
            const int n_rows = 2 * rows_per_band + (band == 0 ? 0 : 1);
            // running vertical max: after an odd c1 row it holds that row (first row of the next pooled
            // row), after an even row the max of rows (2p-1, 2p)
            float state[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) state[k] = -INFINITY;
            for (int r = 0; r < n_rows; ++r) {
                const int oh = r_first + r;
                mb_wait(&t_full[acc], acc_phase);
                tc_after();
                const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 128);
#pragma unroll 1
                for (int cc = 0; cc < 64; cc += 32) {
                    uint32_t r0[32], r1[32];
                    ld32(t0 + cc, r0);
                    if (x3) ld32(t0 + 64 + cc, r1);
                    ld_wait();
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
                tc_before();
                __syncwarp();
                if (lane == 0) mb_arrive(&t_empty[acc]);
                asm volatile("bar.sync 1, 128;" ::: "memory");       // c1 row complete in rowbuf
                // horizontal 3-max for pooled column q, 32 channels
                float hp[32];
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4* rb = rowbuf + ((chh >> 2) + k4) * 128;
                    float4 m = rb[2 * q];
                    const float4 b = rb[2 * q + 1];
                    m.x = fmaxf(m.x, b.x); m.y = fmaxf(m.y, b.y); m.z = fmaxf(m.z, b.z); m.w = fmaxf(m.w, b.w);
                    if (q > 0) {
                        const float4 a = rb[2 * q - 1];
                        m.x = fmaxf(m.x, a.x); m.y = fmaxf(m.y, a.y); m.z = fmaxf(m.z, a.z); m.w = fmaxf(m.w, a.w);
                    }
                    hp[k4 * 4] = m.x; hp[k4 * 4 + 1] = m.y; hp[k4 * 4 + 2] = m.z; hp[k4 * 4 + 3] = m.w;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");       // rowbuf free for the next c1 row
                if ((oh & 1) == 0) {                                  // row 2p: state = max(row 2p-1, row 2p)
#pragma unroll
                    for (int k = 0; k < 32; ++k) state[k] = fmaxf(state[k], hp[k]);
                } else {                                              // row 2p+1 closes pooled row p
                    const int p = oh >> 1;
                    if (p >= p0) {                                    // (the band's extra leading row only primes carry)
                        uint32_t oh_[16], ol_[16];
#pragma unroll
                        for (int k = 0; k < 32; k += 2) {
                            const float a = fminf(fmaxf(state[k], hp[k]), 65504.f), b = fminf(fmaxf(state[k + 1], hp[k + 1]), 65504.f);
                            const __half2 h = __floats2half2_rn(a, b);
                            const float2 hf = __half22float2(h);
                            oh_[k >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                            ol_[k >> 1] = pack2((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
                        }
                        const size_t off = (((size_t)img * 64 + p) * 64 + q) * 64 + chh;
                        uint4* dh = reinterpret_cast<uint4*>(P.out_hi + off);
                        uint4* dl = reinterpret_cast<uint4*>(P.out_lo + off);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            dh[u] = make_uint4(oh_[4 * u], oh_[4 * u + 1], oh_[4 * u + 2], oh_[4 * u + 3]);
                            dl[u] = make_uint4(ol_[4 * u], ol_[4 * u + 1], ol_[4 * u + 2], ol_[4 * u + 3]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 32; ++k) state[k] = hp[k];
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

// host: pack the [64][7][7][4] stem weight into the smem image (hi plane, lo plane)
int stem_tc_pack(ivosw_ctx* c, const float* w_ohwi /*[64][196]*/) {
    using namespace stemtc;
    std::vector<__half> img((size_t)2 * B_PLANE / 2, __float2half_rn(0.f));
    for (int n = 0; n < 64; ++n)
        for (int kh = 0; kh < 7; ++kh)
            for (int kw = 0; kw < 7; ++kw)
                for (int ch = 0; ch < 4; ++ch) {
                    const float w = w_ohwi[(size_t)n * 196 + (kh * 7 + kw) * 4 + ch];
                    const int k = kh * 32 + kw * 4 + ch;
                    const size_t off = ((size_t)(k >> 3) * 8 + (n >> 3)) * 64 + (n & 7) * 8 + (k & 7);   // in halves
                    const __half h = __float2half_rn(w);
                    img[off] = h;
                    img[(size_t)B_PLANE / 2 + off] = __float2half_rn((w - __half2float(h)) * 2048.0f);
                }
    if (!c->stem_wpack) IVOSW_CUDA(cudaMalloc(&c->stem_wpack, (size_t)2 * B_PLANE));
    IVOSW_CUDA(cudaMemcpy(c->stem_wpack, img.data(), (size_t)2 * B_PLANE, cudaMemcpyHostToDevice));
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
    P.crop_hi = (const uint4*)c->crop_hi.p; P.crop_lo = (const uint4*)c->crop_lo.p; P.wpack = (const uint4*)c->stem_wpack;
    P.scale = c->stem_scale; P.shift = c->stem_shift;
    P.out_hi = out.hi; P.out_lo = out.lo;
    P.bands = best; P.n_items = B * best; P.terms = terms;
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
