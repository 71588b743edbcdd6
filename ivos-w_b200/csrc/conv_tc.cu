// Tensor-core implicit-GEMM convolution for sm_100a: tcgen05.mma (kind::f16, fp32 accumulation in
// TMEM) fed by TMA, with the eval-mode BatchNorm scale/shift, residual add and ReLU fused in the
// epilogue.  Serves every bottleneck convolution of res2..res5 (1x1, 3x3, stride 1 and 2).
//
//   out[m, n] = act( (sum_k A[m, k] W[n, k]) * scale[n] + shift[n] + residual[m, n] )
//
// fp32-grade arithmetic on fp16 tensor cores ("split-fp16", IVOSW_CONV_TC_FP16X3)
//   every fp32 value x is stored as two fp16 planes  hi = fp16(x),  lo = fp16((x - hi) * 2048)
//   so x = hi + lo/2048 to ~2^-22 relative, and a product is evaluated as three MMA terms
//       D0 += A_hi W_hi                D1 += A_hi W_lo + A_lo W_hi            D = D0 + D1 / 2048
//   (the lo*lo term, 2^-22 relative, is dropped).  Activations are WRITTEN in this split form by the
//   epilogue, so the two planes cost exactly the bytes of one fp32 tensor.  IVOSW_CONV_TC_FP16X1
//   issues only the first term (plain fp16 inputs).
//
// Implicit GEMM without im2col: activations are NHWC; a tile of 128 output pixels is always a set
// of whole image rows (64x64: 2 rows, 32x32: 4, 16x16: 8, 8x8: two images), so for filter tap
// (kh, kw) the A tile is ONE 5-d TMA box over the activation tensor, shifted by the tap offset, with
// the hardware zero-filling the halo (padding).  Stride-2 convolutions use a 5-d view that splits
// H and W into (half, parity) so that the same box shape applies.  K is walked as (tap, 64-channel
// block); each step lands A_hi/A_lo (128 x 64 fp16, 128B-swizzled, K-major) and W_hi/W_lo
// (BN x 64) in one pipeline stage.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner (the whole warp walks the loop,
// elect.sync issues), then the epilogue warps: 16 in the staged variants (output through a shared-memory
// staging buffer and TMA stores; used wherever a layer has at least two waves of tiles), 8 in the direct
// variant (per-thread 16-byte stores; the few-tile res5 layers).  The compute-bound layers run as CTA PAIRS
// (template flag PAIR_, cta_group::2: two CTAs of a cluster share one M = 256 MMA, each staging its own A rows and half
// of every weight plane; see the comment at launch_conv_tc and DESIGN.md 4.3).  Persistent CTAs walk the (m-tile,
// n-tile) list n-fastest; TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs
// of tile i+1.  Measurements behind these choices: DESIGN.md 4.3, profiles/r1_tc_ceiling.md.
#include "tc_common.cuh"

namespace ivosw {

struct TcParams {
    int M, Cout, Cin;
    int num_taps, cin_blocks;     // K blocks = num_taps * cin_blocks ...
    int num_kb;                   // ... unless two GEMMs share the accumulator (kb_split > 0): blocks [0, kb_split) read
    int kb_split;                 //     the first activation tensor (tap 0), blocks [kb_split, num_kb) the second (tap 1)
    int out_hw;                   // output canvas height (AssessNet: OH == OW, the whole canvas is valid)
    int out_wp;                   // output canvas width (a power of two; > 128: a tile is part of one row)
    int valid_h, valid_w;         // the encoder of the VOS backbone works on canvases larger than the feature map
                                  // (30 x 54 on 32 x 64 ...): positions outside valid_h x valid_w are written as zeros,
                                  // so that they act as the zero padding of the next convolution (0 = no masking)
    int tiles_m, tiles_n;
    int relu, terms;              // terms: 3 (split-fp16) or 1
    int dbg;                      // measurement only (IVOSW_TC_DEBUG): 1 = no TMA operand loads, 2 = no MMAs, 4 = no staged epilogue work
    const float* scale;
    const float* shift;
    const __half* res_hi;
    const __half* res_lo;
    __half* out_hi;
    __half* out_lo;
    unsigned long long* sat_count;   // fp16 range guard events (ivosw_conv_saturation_count)
    TcTap taps[9];
};

// STAGED epilogue (the 1x1 "expand" layers, whose output + residual traffic dominates).  Next to the
// operand ring sits one 32 KB staging buffer (hi + lo planes of a 128 x 64 half tile, 128B-swizzled).  Per
// tile the 16 epilogue warps make two passes (column halves): accumulators come out of TMEM, BN + residual
// + ReLU + the hi/lo split happen in registers, the half tile is written to the staging buffer and ONE
// elected thread sends it back with two TMA stores — fully coalesced 128-byte rows instead of per-thread
// 16-byte stores.
// The residual tile (64 KB) travels through the operand ring as one extra block per tile, fetched by TMA
// (coalesced, deep memory-level parallelism) and signalled on its own barrier pair; the epilogue warps copy
// it into registers as soon as it lands — one tile ahead of its use, right after the previous tile's
// epilogue — and hand the slot straight back, so the block never starves the main loop of ring slots.
// (The first version rewrote the residual block in place and held the slot until the output store had
// read it; with 2 or 4 K blocks per tile on a 3-slot ring that serialised every tile on one residual
// round trip.  A version with per-thread residual loads from global memory was 2x slower on res2:
// 32 rows x 16 bytes per warp instruction is the uncoalesced pattern.  profiles/r1_tc_ceiling.md.)
// PAIR_: two CTAs of a cluster share one M = 256 MMA (cta_group::2); a CTA then stages its own 128 A rows and HALF of
// each weight plane (BN / 2 rows): 48 instead of 64 KB per K block, and a fourth ring slot fits.
template <int BN, int STAGES_, bool STAGED_, bool PAIR_ = false>
struct TcSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 2;     // 16 KB
    static constexpr int B_BYTES = (PAIR_ ? BN / 2 : BN) * TC_BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = STAGES_;
    static constexpr int STG_OFF = STAGES * STAGE_BYTES;                 // 1024-byte aligned (128B swizzle atom)
    static constexpr int NSTG = STAGES_ == 2 ? 2 : 1;                    // a 2-slot ring leaves room for a second buffer
    static constexpr int STG_BYTES = STAGED_ ? NSTG * 2 * TC_BM * 64 * 2 : 0;   // 32 KB each
    static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024 /*align slack*/;
};

struct TcMaps {
    CUtensorMap a_hi, a_lo, w_hi, w_lo;       // operands
    CUtensorMap a2_hi, a2_lo;                 // second activation tensor of a fused pair (kb_split > 0)
    CUtensorMap r_hi, r_lo, o_hi, o_lo;       // residual in / output (STAGED epilogue only)
};

template <int BN, int STAGES_, bool STAGED_, bool PAIR_ = false>
__global__ void __launch_bounds__(tc_threads(STAGED_), 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams P) {
    using S = TcSmem<BN, STAGES_, STAGED_, PAIR_>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full_bar = bars;                       // [STAGES]
    uint64_t* empty_bar = bars + S::STAGES;          // [STAGES]
    uint64_t* tfull_bar = bars + 2 * S::STAGES;      // [2]
    uint64_t* tempty_bar = bars + 2 * S::STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);
    // residual blocks of tile t landed: [(t & 1) * RES_BLOCKS + block].  A residual tile is 64 KB (BN = 128): one ring block,
    // or — in the pair variant, whose slots are 48 KB — one block per 64-column half
    constexpr int RES_BLOCKS = PAIR_ ? BN / 64 : 1;
    constexpr int RES_G = (BN / 64) / RES_BLOCKS;      // 64-column halves per block
    uint64_t* res_full = bars + 2 * S::STAGES + 5;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // PAIR_: a work item is a pair of neighbouring M tiles (256 pixels) x one N tile; CTA rank r of the cluster owns the
    // A rows, the TMEM lanes and the epilogue of M tile 2 * (item / tiles_n) + r.  The host guarantees tiles_m is even.
    const uint32_t cta_rank = PAIR_ ? cluster_ctarank() : 0u;
    const int tile_first = PAIR_ ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR_ ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int num_tiles = PAIR_ ? (P.tiles_m >> 1) * P.tiles_n : P.tiles_m * P.tiles_n;
    const int num_kb = P.num_kb;
    const bool x3 = P.terms == 3;
    // (D0, D1) of one tile are 2 * BN columns: two accumulator stages when they fit the 512 columns (the epilogue of tile i
    // then overlaps the MMAs of tile i + 1), one for the 256-wide pair tiles
    constexpr int ACC_STAGES = 4 * BN <= 512 ? 2 : 1;
    constexpr uint32_t TMEM_COLS = ACC_STAGES * 2 * BN;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a_hi); prefetch_tmap(&maps.a_lo); prefetch_tmap(&maps.w_hi); prefetch_tmap(&maps.w_lo);
        for (int i = 0; i < S::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], tc_epi_warps(STAGED_) * (PAIR_ ? 2 : 1));
            for (int r = 0; r < RES_BLOCKS; ++r) mbar_init(&res_full[i * RES_BLOCKS + r], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (PAIR_) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR_) cluster_sync_all();     // the peer's barriers exist before anything is signalled on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch)
    // overlaps the tail of the previous layer's kernel; its outputs are only touched below this point.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // tiles are walked n-fastest: the CTAs that run together share A tiles (and all weights) in L2
    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx_bytes = x3 ? S::STAGE_BYTES : (S::A_BYTES + S::B_BYTES);
            const int pix_per_img = P.out_hw * P.out_wp;
            int t_local = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int nt = tile % P.tiles_n, mt = (tile / P.tiles_n) * (PAIR_ ? 2 : 1) + (int)cta_rank;
                const int m0 = mt * TC_BM;
                const int n_img = m0 / pix_per_img;
                const int rem0 = m0 - n_img * pix_per_img;
                const int h0 = rem0 / P.out_wp, w0 = rem0 - h0 * P.out_wp;     // w0 != 0 only on canvases wider than a tile
                for (int kb = 0; kb < num_kb; ++kb) {
                    int tap, cb;
                    const CUtensorMap* mh = &maps.a_hi;
                    const CUtensorMap* ml = &maps.a_lo;
                    if (P.kb_split > 0) {
                        if (kb < P.kb_split) { tap = 0; cb = kb; }
                        else { tap = 1; cb = kb - P.kb_split; mh = &maps.a2_hi; ml = &maps.a2_lo; }
                    } else { tap = kb / P.cin_blocks; cb = kb - tap * P.cin_blocks; }
                    const TcTap tp = P.taps[tap];
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * S::STAGE_BYTES;
                    if (P.dbg & 1) {          // measurement: MMA + epilogue only
                        if (!PAIR_ || cta_rank == 0) mbar_arrive(&full_bar[stage]);
                        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    const int c0 = tp.c_add + cb * TC_BK;
                    if constexpr (PAIR_) {
                        // both CTAs' boxes complete on the LEADER's barrier (the leader issues the MMAs and arms it for the
                        // bytes of both); this CTA's half of the weight planes: rows [rank * BN / 2, + BN / 2) of the N tile
                        const uint32_t fb = mapa_u32(&full_bar[stage], 0);
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * tx_bytes);
                        const int wrow = nt * BN + (int)cta_rank * (BN / 2);
                        tma_load_5d_pair(st, mh, fb, c0, w0 + tp.w_add, tp.p, h0 + tp.h_add, n_img);
                        tma_load_2d_pair(st + 2 * S::A_BYTES, &maps.w_hi, fb, kb * TC_BK, wrow);
                        if (x3) {
                            tma_load_5d_pair(st + S::A_BYTES, ml, fb, c0, w0 + tp.w_add, tp.p, h0 + tp.h_add, n_img);
                            tma_load_2d_pair(st + 2 * S::A_BYTES + S::B_BYTES, &maps.w_lo, fb, kb * TC_BK, wrow);
                        }
                    } else {
                        mbar_expect_tx(&full_bar[stage], tx_bytes);
                        tma_load_5d(st, mh, &full_bar[stage], c0, w0 + tp.w_add, tp.p, h0 + tp.h_add, n_img);
                        tma_load_2d(st + 2 * S::A_BYTES, &maps.w_hi, &full_bar[stage], kb * TC_BK, nt * BN);
                        if (x3) {
                            tma_load_5d(st + S::A_BYTES, ml, &full_bar[stage], c0, w0 + tp.w_add, tp.p, h0 + tp.h_add, n_img);
                            tma_load_2d(st + 2 * S::A_BYTES + S::B_BYTES, &maps.w_lo, &full_bar[stage], kb * TC_BK, nt * BN);
                        }
                    }
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
                if constexpr (STAGED_) {
                    if (P.res_hi != nullptr) {
                        // residual block of this tile: both column halves, hi/lo planes; completion goes to
                        // res_full — the slot's own full barrier is not involved (the MMA issuer keeps a parity
                        // bit per slot and only counts the uses it waits on)
#pragma unroll
                        for (int r = 0; r < RES_BLOCKS; ++r) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            uint8_t* st = smem + stage * S::STAGE_BYTES;
                            uint64_t* rb = &res_full[(t_local & 1) * RES_BLOCKS + r];
                            mbar_expect_tx(rb, (x3 ? 2u : 1u) * RES_G * TC_BM * 128);
#pragma unroll
                            for (int gg = 0; gg < RES_G; ++gg) {
                                const int g = r * RES_G + gg;
                                tma_load_2d(st + gg * 2 * TC_BM * 128, &maps.r_hi, rb, nt * BN + g * 64, m0);
                                if (x3) tma_load_2d(st + gg * 2 * TC_BM * 128 + TC_BM * 128, &maps.r_lo, rb, nt * BN + g * 64, m0);
                            }
                            if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                    ++t_local;
                }
            }
        }
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        // (PAIR_: the leader CTA issues for both; the peer's warp 1 only owns its TMEM allocation)
        // The whole warp walks the loop with warp-uniform values and elect.sync picks the issuing lane: descriptors
        // and barrier addresses then live in uniform registers and every UTCHMMA / UTCBAR issues straight, without
        // the ELECT / BRA.U.ANY waterfall ptxas wraps around per-thread operands.
        if (!PAIR_ || cta_rank == 0) {
            // instruction descriptor: D = F32, A = B = F16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            const uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            // pair form: M = 256 over the two CTAs, N = BN; each CTA supplies BN / 2 rows of the B operand
            const uint32_t idesc_pair = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
            (void)idesc_pair;
            const bool no_mma = (P.dbg & 2) != 0;                // measurement: loads + epilogue only
            const bool skip_res_slot = STAGED_ && P.res_hi != nullptr;
            const uint32_t smem_base = smem_u32(smem);
            int stage = 0; uint32_t full_bits = 0;               // parity of the next operand block, per slot
            int acc = 0; uint32_t acc_phase = 0;
            // Between two bursts the tensor pipe only has the last instruction of the previous burst queued (64-128
            // cycles of work): everything the issuing thread does from "block landed" to the first MMA of the next burst
            // is a pipe bubble (scripts/umma_rate2.cu: 896 cycles per K block with a wait between bursts, 775 free
            // running).  So (1) the NEXT block's barrier is probed, without blocking, before the last K step of the
            // current burst — when the loads keep up, the next burst starts without any wait in between (797 cycles);
            // (2) the last K step issues its N = BN instruction first and the N = 2 BN one last; (3) descriptors are
            // one add away from a per-kernel constant.
            const uint64_t desc0 = make_sw128_desc(smem_base);   // A_hi of slot 0; everything else is + constant
            constexpr uint32_t SLOT16 = S::STAGE_BYTES >> 4, ALO16 = S::A_BYTES >> 4, B16 = (2 * S::A_BYTES) >> 4;
            bool ready = false;                                  // the probe already saw the next block
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);      // epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(acc * 2 * BN);
                const uint32_t d1 = d0 + BN;
                const bool last_tile = tile + tile_step >= num_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (!ready) mbar_wait(&full_bar[stage], (full_bits >> stage) & 1u);
                    full_bits ^= 1u << stage;
                    tc_fence_after();
                    const uint64_t a_hi = desc0 + (uint64_t)((uint32_t)stage * SLOT16);
                    const uint64_t a_lo = a_hi + ALO16, b_hi = a_hi + B16;
                    const bool last_kb = kb == num_kb - 1;
                    if (!no_mma && elect_one()) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 16 - 1; ++k) {
                            const uint64_t adv = (uint64_t)(2 * k);                  // +32 bytes inside the swizzle row
                            const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
                            if constexpr (PAIR_) {
                                // three N = BN pair instructions in the accumulation order of the single-CTA form
                                const uint64_t b_lo = b_hi + (uint64_t)(S::B_BYTES >> 4);
                                umma_f16_pair(d0, a_hi + adv, b_hi + adv, idesc_pair, accum);
                                if (x3) {
                                    umma_f16_pair(d1, a_hi + adv, b_lo + adv, idesc_pair, accum);
                                    umma_f16_pair(d1, a_lo + adv, b_hi + adv, idesc_pair, 1u);
                                }
                            } else if (x3) {
                                // [D0 | D1] += A_hi * [W_hi ; W_lo]^T as ONE N = 2*BN instruction (the two weight
                                // tiles are contiguous in the stage, the two accumulators contiguous in TMEM):
                                // A_hi is read from shared memory once instead of twice
                                umma_f16(d0, a_hi + adv, b_hi + adv, idesc_2n, accum);
                                umma_f16(d1, a_lo + adv, b_hi + adv, idesc, 1u);
                            } else {
                                umma_f16(d0, a_hi + adv, b_hi + adv, idesc, accum);
                            }
                        }
                    }
                    __syncwarp();
                    // where the next operand block sits: the slot after this one, one further at the end of a tile
                    // that carries a residual block
                    int nstage = stage + 1 == S::STAGES ? 0 : stage + 1;
                    if (last_kb && skip_res_slot) nstage = (nstage + RES_BLOCKS) % S::STAGES;
                    ready = !(last_kb && last_tile) && mbar_test(&full_bar[nstage], (full_bits >> nstage) & 1u);
                    if (elect_one()) {
                        if (!no_mma) {
                            const uint64_t adv = (uint64_t)(2 * (TC_BK / 16 - 1));
                            if constexpr (PAIR_) {
                                const uint64_t b_lo = b_hi + (uint64_t)(S::B_BYTES >> 4);
                                if (x3) umma_f16_pair(d1, a_lo + adv, b_hi + adv, idesc_pair, 1u);
                                umma_f16_pair(d0, a_hi + adv, b_hi + adv, idesc_pair, 1u);
                                if (x3) umma_f16_pair(d1, a_hi + adv, b_lo + adv, idesc_pair, 1u);
                            } else if (x3) {
                                umma_f16(d1, a_lo + adv, b_hi + adv, idesc, 1u);
                                umma_f16(d0, a_hi + adv, b_hi + adv, idesc_2n, 1u);
                            } else {
                                umma_f16(d0, a_hi + adv, b_hi + adv, idesc, 1u);
                            }
                        }
                        if constexpr (PAIR_) {                   // both CTAs' slots / both CTAs' epilogues
                            umma_commit_pair(&empty_bar[stage]);
                            if (last_kb) umma_commit_pair(&tfull_bar[acc]);
                        } else {
                            umma_commit(&empty_bar[stage]);      // smem slot reusable once these MMAs retire
                            if (last_kb) umma_commit(&tfull_bar[acc]);
                        }
                    }
                    __syncwarp();
                    stage = nstage;
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ======================================= epilogue =======================================
        const int e = warp - 2;                 // 0..7
        const int quarter = warp & 3;           // TMEM lanes this warp may touch: 32*quarter .. +31
        const int half = e >> 2;                // which half of the tile's columns
        constexpr int COLS = BN / 2;            // columns per warp
        const int row = quarter * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        if constexpr (STAGED_) {
            constexpr int NP = BN / 64;                         // passes of 64 columns per tile
            // 16 warps: quarter = TMEM lane quarter (a hardware rule: warp % 4), csub = which 16-column slice
            // of the 64-column pass this warp owns
            const int csub = e >> 2;
            const bool leader = threadIdx.x == 64;
            const bool has_res = P.res_hi != nullptr;
            const bool relu = P.relu != 0;
            // (with two staging buffers — the K = 64 layers on a 2-slot ring — pass g owns buffer g and only waits
            //  for its own store of the previous tile)
            uint32_t soff[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) soff[q] = (uint32_t)row * 128u + (uint32_t)(((2 * csub + q) ^ (row & 7)) << 4);
            // (The residual tile is copied into registers as soon as it lands and its ring slot handed straight back.  A
            //  variant that left it in the slot and read it pass by pass — no 32 extra registers, no spills, the form
            //  conv_stack.cu uses — was measured here too: equal in the bench, but slower per layer under ncu
            //  (res2 conv3 268 vs 189 us): with 2-3 ring slots a residual block that stays until the tile's last pass keeps
            //  the next tile's residual from being fetched during this tile's epilogue.)
            float rf[NP][16];                                   // this thread's residual values: [pass][column]
#pragma unroll
            for (int g = 0; g < NP; ++g)
#pragma unroll
                for (int k = 0; k < 16; ++k) rf[g][k] = 0.f;
            int stage = 0, t_local = 0;                         // ring position of the residual blocks
            __half2 sat = __float2half2_rn(0.f);            // running max of |hi| (fp16 range guard)
            const bool masking = P.valid_w > 0;                 // kernel-uniform: canvases larger than the feature map
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int nt = tile % P.tiles_n, mt = (tile / P.tiles_n) * (PAIR_ ? 2 : 1) + (int)cta_rank;
                bool row_valid = true;
                if (masking) {                                  // this thread's output pixel inside the feature map?
                    const int rem = (mt * TC_BM + row) % (P.out_hw * P.out_wp);
                    row_valid = (rem / P.out_wp) < P.valid_h && (rem % P.out_wp) < P.valid_w;
                }
                if (has_res) {
                    // residual tile -> registers (every epilogue thread waits on every use of res_full, in order),
                    // then the slot goes straight back to the producer
                    stage = (stage + num_kb) % S::STAGES;
#pragma unroll
                    for (int r = 0; r < RES_BLOCKS; ++r) {
                    mbar_wait(&res_full[(t_local & 1) * RES_BLOCKS + r], (uint32_t)(t_local >> 1) & 1u);
                    const uint32_t rs = smem_u32(smem + stage * S::STAGE_BYTES);
#pragma unroll
                    for (int gg = 0; gg < RES_G; ++gg)
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int g = r * RES_G + gg;
                            const uint4 h4 = lds128(rs + gg * 2 * TC_BM * 128 + soff[q]);
                            const uint4 l4 = x3 ? lds128(rs + gg * 2 * TC_BM * 128 + TC_BM * 128 + soff[q]) : make_uint4(0, 0, 0, 0);
                            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[u]));
                                const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[u]));
                                rf[g][q * 8 + u * 2 + 0] = fmaf(lf.x, 1.0f / 2048.0f, hf.x);
                                rf[g][q * 8 + u * 2 + 1] = fmaf(lf.y, 1.0f / 2048.0f, hf.y);
                            }
                        }
                    // The slot is about to be overwritten by TMA (async proxy): the generic-proxy reads above must be
                    // complete first.  A barrier alone orders them against other threads' generic accesses only —
                    // releasing right after it let the next block land under still-queued reads.
                    fence_proxy_async();
                    group_bar(3, 512);
                    if (leader) mbar_arrive(&empty_bar[stage]);
                    stage = (stage + 1) % S::STAGES;
                    }
                    ++t_local;
                }
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                if (P.dbg & 4) {                                // measurement: main loop only
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if constexpr (PAIR_) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0)); else mbar_arrive(&tempty_bar[acc]); }
                    if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
                    continue;
                }
#pragma unroll
                for (int g = 0; g < NP; ++g) {
                    const int n = nt * BN + g * 64 + csub * 16;                 // first output channel of this thread
                    const uint32_t t_d0 = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                          (uint32_t)(acc * 2 * BN + g * 64 + csub * 16);
                    uint32_t r0[16], r1[16];
                    tmem_ld16(t_d0, r0);
                    if (x3) tmem_ld16(t_d0 + BN, r1);
                    tmem_ld_wait();
                    if (g == NP - 1) {                          // last TMEM read of the tile: hand the accumulator back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if constexpr (PAIR_) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0)); else mbar_arrive(&tempty_bar[acc]); }
                    }
                    uint32_t oh[8], ol[8];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 s0 = __ldg(reinterpret_cast<const float4*>(P.scale + n + q * 8));
                        const float4 s1 = __ldg(reinterpret_cast<const float4*>(P.scale + n + q * 8 + 4));
                        const float4 h0 = __ldg(reinterpret_cast<const float4*>(P.shift + n + q * 8));
                        const float4 h1 = __ldg(reinterpret_cast<const float4*>(P.shift + n + q * 8 + 4));
                        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                        const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                        float v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            float a = __uint_as_float(r0[q * 8 + k]);
                            if (x3) a = fmaf(__uint_as_float(r1[q * 8 + k]), 1.0f / 2048.0f, a);
                            v[k] = fmaf(a, sc[k], sh[k]);
                        }
                        if (has_res) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] += rf[g][q * 8 + k];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float a = v[u * 2], b = v[u * 2 + 1];
                            if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                            else { a = fmaxf(a, -65504.f); b = fmaxf(b, -65504.f); }
                            a = fminf(a, 65504.f);                                      // fp16 range guard
                            b = fminf(b, 65504.f);
                            const __half2 h = __floats2half2_rn(a, b);
                            const float2 hf = __half22float2(h);
                            oh[q * 4 + u] = *reinterpret_cast<const uint32_t*>(&h);
                            sat = __hmax2(sat, __habs2(*reinterpret_cast<const __half2*>(&oh[q * 4 + u])));
                            ol[q * 4 + u] = pack_half2((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
                        }
                    }
                    if (masking && !row_valid) {            // outside the feature map: zeros = the next layer's padding
#pragma unroll
                        for (int u = 0; u < 8; ++u) { oh[u] = 0u; ol[u] = 0u; }
                    }
                    uint8_t* stg = smem + S::STG_OFF + (S::NSTG == 2 ? g * 2 * TC_BM * 128 : 0);
                    const uint32_t stg_hi = smem_u32(stg), stg_lo = stg_hi + TC_BM * 128;
                    if (leader) {                           // the previous stores from this buffer have read it
                        if (S::NSTG == 2) bulk_wait_read1(); else bulk_wait_read0();
                    }
                    group_bar(2, 512);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        sts128(stg_hi + soff[q], make_uint4(oh[q * 4], oh[q * 4 + 1], oh[q * 4 + 2], oh[q * 4 + 3]));
                        sts128(stg_lo + soff[q], make_uint4(ol[q * 4], ol[q * 4 + 1], ol[q * 4 + 2], ol[q * 4 + 3]));
                    }
                    fence_proxy_async();                    // generic-proxy smem writes -> visible to the TMA store
                    group_bar(1, 512);                      // half tile staged
                    if (leader) {
                        tma_store_2d(&maps.o_hi, stg, nt * BN + g * 64, mt * TC_BM);
                        tma_store_2d(&maps.o_lo, stg + TC_BM * 128, nt * BN + g * 64, mt * TC_BM);
                        bulk_commit();
                    }
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
            if (leader) bulk_wait0();
            if (sat_hit(sat)) atomicAdd(P.sat_count, 1ull);
        } else {
            __half2 sat = __float2half2_rn(0.f);            // running max of |hi| (fp16 range guard)
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int nt = tile % P.tiles_n, mt = (tile / P.tiles_n) * (PAIR_ ? 2 : 1) + (int)cta_rank;
                const long long m = (long long)mt * TC_BM + row;
                const bool row_ok = m < P.M;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_d0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * BN + half * COLS);
#pragma unroll 1
                for (int cc = 0; cc < ((P.dbg & 4) ? 0 : COLS); cc += 32) {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(t_d0 + cc, r0);
                    if (x3) tmem_ld32(t_d0 + BN + cc, r1);
                    tmem_ld_wait();
                    const int n = nt * BN + half * COLS + cc;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float a = __uint_as_float(r0[j]);
                        if (x3) a = fmaf(__uint_as_float(r1[j]), 1.0f / 2048.0f, a);
                        v[j] = fmaf(a, __ldg(P.scale + n + j), __ldg(P.shift + n + j));
                    }
                    if (row_ok) {
                        const size_t off = (size_t)m * P.Cout + n;
                        if (P.res_hi) {
                            const uint4* rh = reinterpret_cast<const uint4*>(P.res_hi + off);
                            const uint4* rl = reinterpret_cast<const uint4*>(P.res_lo + off);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 h4 = __ldg(rh + q);
                                const uint4 l4 = x3 ? __ldg(rl + q) : make_uint4(0, 0, 0, 0);
                                const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[u]));
                                    const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[u]));
                                    v[q * 8 + u * 2 + 0] += fmaf(lf.x, 1.0f / 2048.0f, hf.x);
                                    v[q * 8 + u * 2 + 1] += fmaf(lf.y, 1.0f / 2048.0f, hf.y);
                                }
                            }
                        }
                        uint32_t oh[16], ol[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float a = v[j], b = v[j + 1];
                            if (P.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                            a = fminf(fmaxf(a, -65504.f), 65504.f);          // fp16 range guard
                            b = fminf(fmaxf(b, -65504.f), 65504.f);
                            const __half2 h = __floats2half2_rn(a, b);
                            const float2 hf = __half22float2(h);
                            oh[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                            sat = __hmax2(sat, __habs2(*reinterpret_cast<const __half2*>(&oh[j >> 1])));
                            ol[j >> 1] = pack_half2((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
                        }
                        uint4* ph = reinterpret_cast<uint4*>(P.out_hi + off);
                        uint4* pl = reinterpret_cast<uint4*>(P.out_lo + off);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if ((P.dbg & 32) && oh[q * 4] != 0x7fff7fffu) continue;     // measurement: compute, no stores
                            ph[q] = make_uint4(oh[q * 4], oh[q * 4 + 1], oh[q * 4 + 2], oh[q * 4 + 3]);
                            pl[q] = make_uint4(ol[q * 4], ol[q * 4 + 1], ol[q * 4 + 2], ol[q * 4 + 3]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if constexpr (PAIR_) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0)); else mbar_arrive(&tempty_bar[acc]); }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
            if (sat_hit(sat)) atomicAdd(P.sat_count, 1ull);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR_) cluster_sync_all();     // no CTA leaves (or frees TMEM) while its peer can still signal or be signalled
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR_)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point unavailable"); return IVOSW_ERR_CUDA; }
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
        set_error(buf);
        return IVOSW_ERR_CUDA;
    }
    return IVOSW_OK;
}

// 5-d view of an NHWC fp16 activation plane whose boxes are "whole rows of the OUTPUT grid".
int encode_act_map(CUtensorMap* map, const __half* base, int B, int H, int C, int stride, int out_hw) {
    const int W = H;
    const int Wb = out_hw;
    const int Hb = (TC_BM / Wb) < out_hw ? (TC_BM / Wb) : out_hw;
    const int Nb = TC_BM / (Wb * Hb);
    // (an 8x8 tile spans two images: with B == 1 the map claims a second image; the workspace behind it is
    //  allocated and the rows it produces are never stored)
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)Wb, 1, (cuuint32_t)Hb, (cuuint32_t)Nb};
    const cuuint64_t e = sizeof(__half);
    if (stride == 1) {            // (C, W, 1, H, N)
        dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = B > Nb ? B : Nb;
        strides[0] = (cuuint64_t)C * e; strides[1] = (cuuint64_t)W * C * e; strides[2] = (cuuint64_t)W * C * e;
        strides[3] = (cuuint64_t)H * W * C * e;
    } else {                      // (2C [w parity, c], W/2, 2 [h parity], H/2, N)
        dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B > Nb ? B : Nb;
        strides[0] = (cuuint64_t)2 * C * e; strides[1] = (cuuint64_t)W * C * e; strides[2] = (cuuint64_t)2 * W * C * e;
        strides[3] = (cuuint64_t)H * W * C * e;
    }
    return encode_map(map, base, 5, dims, strides, box);
}

// The same for rectangular power-of-two canvases (Hp x Wp): a 128-pixel tile is Hb rows of Wb = min(Wp_out, 128) pixels.
int encode_act_map_g(CUtensorMap* map, const __half* base, int B, int Hp, int Wp, int C, int stride, int out_hp, int out_wp) {
    const int Wb = out_wp < TC_BM ? out_wp : TC_BM;
    const int Hb = (TC_BM / Wb) < out_hp ? (TC_BM / Wb) : out_hp;
    const int Nb = TC_BM / (Wb * Hb);
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)Wb, 1, (cuuint32_t)Hb, (cuuint32_t)Nb};
    const cuuint64_t e = sizeof(__half);
    if (stride == 1) {
        dims[0] = C; dims[1] = Wp; dims[2] = 1; dims[3] = Hp; dims[4] = B > Nb ? B : Nb;
        strides[0] = (cuuint64_t)C * e; strides[1] = (cuuint64_t)Wp * C * e; strides[2] = (cuuint64_t)Wp * C * e;
        strides[3] = (cuuint64_t)Hp * Wp * C * e;
    } else {
        dims[0] = 2 * C; dims[1] = Wp / 2; dims[2] = 2; dims[3] = Hp / 2; dims[4] = B > Nb ? B : Nb;
        strides[0] = (cuuint64_t)2 * C * e; strides[1] = (cuuint64_t)Wp * C * e; strides[2] = (cuuint64_t)2 * Wp * C * e;
        strides[3] = (cuuint64_t)Hp * Wp * C * e;
    }
    return encode_map(map, base, 5, dims, strides, box);
}

// output / residual tile of a tensor whose rows are `ld` channels apart (a channel slice of a concatenation buffer)
int encode_out_map_ld(CUtensorMap* map, const __half* base, long long M, int Cout, int ld) {
    cuuint64_t dims[2] = {(cuuint64_t)Cout, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__half)};
    cuuint32_t box[2] = {64, (cuuint32_t)TC_BM};
    return encode_map(map, base, 2, dims, strides, box);
}

int encode_w_map(CUtensorMap* map, const __half* base, int K, int Cout, int BN) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
    return encode_map(map, base, 2, dims, strides, box);
}

int encode_out_map(CUtensorMap* map, const __half* base, long long M, int Cout) {
    cuuint64_t dims[2] = {(cuuint64_t)Cout, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)Cout * sizeof(__half)};
    cuuint32_t box[2] = {64, (cuuint32_t)TC_BM};
    return encode_map(map, base, 2, dims, strides, box);
}

template <int BN, int STAGES, bool STAGED, bool PAIR = false>
static int launch_tc_variant(ivosw_ctx* c, const TcMaps& maps, const TcParams& P, cudaStream_t s) {
    using S = TcSmem<BN, STAGES, STAGED, PAIR>;
    static_assert(S::TOTAL <= 232448, "shared memory budget (227 KB)");
    // the opt-in above 48 KB of dynamic shared memory is a per-DEVICE function attribute: one Engine per device may
    // live in the same process (engine.get_engine), so remember it per device, not per process
    static bool attr[64] = {};
    const int dv = c->device & 63;
    if (!attr[dv]) {
        IVOSW_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, STAGED, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        S::TOTAL));
        attr[dv] = true;
    }
    // PAIR: clusters of two CTAs (an SM pair of one TPC), one cluster per work item of two M tiles
    const int tiles = PAIR ? (P.tiles_m / 2) * P.tiles_n : P.tiles_m * P.tiles_n;
    const int slots = PAIR ? c->sm_count / 2 : c->sm_count;
    const int grid = (tiles < slots ? tiles : slots) * (PAIR ? 2 : 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc_threads(STAGED)); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = s;
    cudaLaunchAttribute lattr[2];
    int na = 0;
    static const bool pdl = !(getenv("IVOSW_PDL") && atoi(getenv("IVOSW_PDL")) == 0);
    if (pdl) {
        lattr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        lattr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (PAIR) {
        lattr[na].id = cudaLaunchAttributeClusterDimension;
        lattr[na].val.clusterDim.x = 2; lattr[na].val.clusterDim.y = 1; lattr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = lattr; cfg.numAttrs = na;
    IVOSW_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, STAGES, STAGED, PAIR>, maps, P));
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

static void fill_tap(TcTap& t, int oy, int ox, int stride, int cin) {
    // input offset (oy, ox) of a tap relative to stride * out
    if (stride == 1) {
        t.c_add = 0; t.w_add = ox; t.p = 0; t.h_add = oy;
    } else {                                        // ih = 2*oh + oy = 2*(oh + d) + parity
        const int py = oy & 1, px = ox & 1;
        t.p = py; t.h_add = (oy - py) / 2;
        t.c_add = px * cin; t.w_add = (ox - px) / 2;
    }
}

// One convolution layer — or, with `fuse` set, the tail of a bottleneck's first block as ONE GEMM:
//     out = relu( bn3(conv3(in)) + bn_d(downsample(in2)) )  =  relu( [s3 W3 | sd Wd] [in ; in2] + (b3 + bd) )
// (both 1x1; the BatchNorm scales are folded into the concatenated weight fuse->w_hi/w_lo, the shifts added; the
// downsample output never exists in memory: one 537 MB write and one 537 MB read less in res2.0 alone).
int launch_conv_tc(ivosw_ctx* c, const ConvLayer& L, const SplitAct& in, const SplitAct* residual, const SplitAct& out,
                   int B, int terms, cudaStream_t s, const FusedTail* fuse, const SplitAct* in2) {
    // 128-column tiles, unless they would leave more than half of the SMs without a tile (small unit batches: an 8-GPU
    // shard, the chunks of the host-buffer path): then 64-column tiles double the CTAs at work
    const long long m_tiles = ((long long)B * L.out_hw * L.out_hw + TC_BM - 1) / TC_BM;
    static const int narrow_below = getenv("IVOSW_NARROW_BELOW") ? atoi(getenv("IVOSW_NARROW_BELOW")) : -1;
    const long long narrow_limit = narrow_below >= 0 ? narrow_below : c->sm_count / 2;
    const int BN = (L.cout >= 128 && m_tiles * (L.cout / 128) > narrow_limit) ? 128 : 64;
    const int K = fuse ? fuse->k_total : L.k * L.k * L.cin;
    // Epilogue through the staging buffer + TMA stores: always for the 1x1 "expand" layers (conv3 and downsample:
    // Cout = 4 * planes, the most output / residual bytes), and for every other layer with at least two waves of tiles
    // (coalesced stores beat the per-thread 16-byte ones by 3-40 us per layer; with fewer tiles the two-pass epilogue of
    // the last tile is a longer tail than it saves)
    const long long n_tiles = m_tiles * (L.cout / BN);
    const bool staged = ((L.k == 1 && L.cout >= 256 && (residual != nullptr || L.is_downsample || fuse)) || n_tiles >= 2 * c->sm_count) &&
                        getenv("IVOSW_NO_STAGED_EPILOGUE") == nullptr;
    int rc;
    TcMaps maps;
    memset(&maps, 0, sizeof maps);
    TcParams P;
    memset(&P, 0, sizeof P);
    P.M = B * L.out_hw * L.out_hw;
    if ((rc = encode_act_map(&maps.a_hi, in.hi, B, L.in_hw, L.cin, L.stride, L.out_hw))) return rc;
    if ((rc = encode_act_map(&maps.a_lo, in.lo, B, L.in_hw, L.cin, L.stride, L.out_hw))) return rc;
    if (fuse) {
        if ((rc = encode_act_map(&maps.a2_hi, in2->hi, B, fuse->in_hw2, fuse->cin2, fuse->stride2, L.out_hw))) return rc;
        if ((rc = encode_act_map(&maps.a2_lo, in2->lo, B, fuse->in_hw2, fuse->cin2, fuse->stride2, L.out_hw))) return rc;
    }
    // CTA pairs (cta_group::2, template flag PAIR_): the compute-bound layers, when the number of M tiles is even —
    //   * conv1 / conv2 (no residual) with at least eight K blocks (res3.0.conv1 with four is HBM-bound: 150 vs 127 us),
    //   * the fused conv3 + downsample GEMMs of res4.0 / res5.0 (12 and 24 K blocks: 132 -> 120 and 125 -> 110 us).
    // A CTA then ingests 48 instead of 64 KB per K block — one SM takes in ~79 B/clk and 64 KB per 768 tensor cycles is
    // more than that (scripts/umma_rate2.cu mode 14) — and the ring gets a fourth slot.  The HBM-bound expand layers stay
    // single-CTA: as pairs two CTAs walk one residual + output stream in lock step (res3.1.conv3 150 vs 108 us,
    // res4.x.conv3 90 vs 75 us).  IVOSW_PAIR=0: off; IVOSW_PAIR=2: the expand layers too (measurement).
    static const int pair_mode = getenv("IVOSW_PAIR") ? atoi(getenv("IVOSW_PAIR")) : 1;
    const bool pair_ok = pair_mode != 0;
    const bool expand = residual != nullptr || fuse != nullptr || L.is_downsample;
    const bool want = expand ? (pair_mode == 2 || (fuse != nullptr && K / TC_BK >= 12)) : K / TC_BK >= 8;
    const bool pair = pair_ok && BN == 128 && K / TC_BK >= 2 && (m_tiles % 2) == 0 &&
                      (long long)B * L.out_hw * L.out_hw == m_tiles * TC_BM && want;
    if ((rc = encode_w_map(&maps.w_hi, fuse ? fuse->w_hi : L.w_hi, K, L.cout, pair ? BN / 2 : BN))) return rc;
    if ((rc = encode_w_map(&maps.w_lo, fuse ? fuse->w_lo : L.w_lo, K, L.cout, pair ? BN / 2 : BN))) return rc;
    if (staged) {
        if ((rc = encode_out_map(&maps.o_hi, out.hi, P.M, L.cout))) return rc;
        if ((rc = encode_out_map(&maps.o_lo, out.lo, P.M, L.cout))) return rc;
        if (residual) {
            if ((rc = encode_out_map(&maps.r_hi, residual->hi, P.M, L.cout))) return rc;
            if ((rc = encode_out_map(&maps.r_lo, residual->lo, P.M, L.cout))) return rc;
        }
    }
    P.Cout = L.cout; P.Cin = L.cin;
    P.num_taps = L.k * L.k; P.cin_blocks = L.cin / TC_BK;
    P.num_kb = K / TC_BK; P.kb_split = fuse ? L.cin / TC_BK : 0;
    P.out_hw = L.out_hw; P.out_wp = L.out_hw;
    P.tiles_m = (P.M + TC_BM - 1) / TC_BM; P.tiles_n = L.cout / BN;
    P.relu = L.relu ? 1 : 0; P.terms = terms;
    { const char* e = getenv("IVOSW_TC_DEBUG"); P.dbg = e ? atoi(e) : 0; }
    P.scale = fuse ? fuse->scale : L.scale; P.shift = fuse ? fuse->shift : L.shift;
    P.res_hi = residual ? residual->hi : nullptr; P.res_lo = residual ? residual->lo : nullptr;
    P.out_hi = out.hi; P.out_lo = out.lo;
    P.sat_count = c->sat_count;
    if (fuse) {
        fill_tap(P.taps[0], 0, 0, 1, L.cin);
        fill_tap(P.taps[1], 0, 0, fuse->stride2, fuse->cin2);
    } else {
        for (int kh = 0; kh < L.k; ++kh)
            for (int kw = 0; kw < L.k; ++kw) fill_tap(P.taps[kh * L.k + kw], kh - L.pad, kw - L.pad, L.stride, L.cin);
    }
    if (staged) {
        if (BN == 64) return launch_tc_variant<64, 4, true>(c, maps, P, s);
        // one K block per tile (res2: Cin = 64): output/residual traffic is everything, two staging buffers
        // matter more than a third ring slot
        if (P.num_kb == 1 && getenv("IVOSW_NO_STAGED2") == nullptr) return launch_tc_variant<128, 2, true>(c, maps, P, s);
        if (pair) return launch_tc_variant<128, 4, true, true>(c, maps, P, s);
        return launch_tc_variant<128, 3, true>(c, maps, P, s);
    }
    if (pair) return launch_tc_variant<128, 4, false, true>(c, maps, P, s);
    return BN == 128 ? launch_tc_variant<128, 3, false>(c, maps, P, s) : launch_tc_variant<64, 4, false>(c, maps, P, s);
}

// One convolution of the VOS encoder (manet_encoder.cu): rectangular power-of-two canvases around a feature map of any
// size, dilation, stride 1 or 2, output into a channel slice, optional fused second 1x1 GEMM (the downsample branch)
// and residual.  Same kernel, staged epilogue, positions outside the feature map written as zeros.
int launch_conv_tc_g(ivosw_ctx* c, const GConv& L, const SplitAct& in, const SplitAct* in2, const SplitAct* residual,
                     const SplitAct& out, int out_ld, int B, int terms, cudaStream_t s) {
    const int BN = L.cout >= 128 ? 128 : 64;
    const int k3 = L.k * L.k * L.cin;
    const int K = k3 + L.cin2;
    int rc;
    TcMaps maps;
    memset(&maps, 0, sizeof maps);
    TcParams P;
    memset(&P, 0, sizeof P);
    P.M = B * L.out_hp * L.out_wp;
    if ((rc = encode_act_map_g(&maps.a_hi, in.hi, B, L.in_hp, L.in_wp, L.cin, L.stride, L.out_hp, L.out_wp))) return rc;
    if ((rc = encode_act_map_g(&maps.a_lo, in.lo, B, L.in_hp, L.in_wp, L.cin, L.stride, L.out_hp, L.out_wp))) return rc;
    if (L.cin2) {
        if ((rc = encode_act_map_g(&maps.a2_hi, in2->hi, B, L.in2_hp, L.in2_wp, L.cin2, L.stride2, L.out_hp, L.out_wp))) return rc;
        if ((rc = encode_act_map_g(&maps.a2_lo, in2->lo, B, L.in2_hp, L.in2_wp, L.cin2, L.stride2, L.out_hp, L.out_wp))) return rc;
    }
    // CTA pairs for the compute-bound layers, as in launch_conv_tc
    static const bool pair_ok = !(getenv("IVOSW_PAIR") && atoi(getenv("IVOSW_PAIR")) == 0);
    const bool pair = pair_ok && BN == 128 && residual == nullptr && L.cin2 == 0 && K / TC_BK >= 8 && (P.M % (2 * TC_BM)) == 0;
    if ((rc = encode_w_map(&maps.w_hi, L.w_hi, K, L.cout, pair ? BN / 2 : BN))) return rc;
    if ((rc = encode_w_map(&maps.w_lo, L.w_lo, K, L.cout, pair ? BN / 2 : BN))) return rc;
    if ((rc = encode_out_map_ld(&maps.o_hi, out.hi, P.M, L.cout, out_ld))) return rc;
    if ((rc = encode_out_map_ld(&maps.o_lo, out.lo, P.M, L.cout, out_ld))) return rc;
    if (residual) {
        if ((rc = encode_out_map_ld(&maps.r_hi, residual->hi, P.M, L.cout, L.cout))) return rc;
        if ((rc = encode_out_map_ld(&maps.r_lo, residual->lo, P.M, L.cout, L.cout))) return rc;
    }
    P.Cout = L.cout; P.Cin = L.cin;
    P.num_taps = L.k * L.k; P.cin_blocks = L.cin / TC_BK;
    P.num_kb = K / TC_BK; P.kb_split = L.cin2 ? k3 / TC_BK : 0;
    P.out_hw = L.out_hp; P.out_wp = L.out_wp; P.valid_h = L.valid_h; P.valid_w = L.valid_w;
    P.tiles_m = (P.M + TC_BM - 1) / TC_BM; P.tiles_n = L.cout / BN;
    P.relu = L.relu; P.terms = terms;
    P.scale = L.scale; P.shift = L.shift;
    P.res_hi = residual ? residual->hi : nullptr; P.res_lo = residual ? residual->lo : nullptr;
    P.out_hi = out.hi; P.out_lo = out.lo;
    P.sat_count = c->sat_count;
    if (L.cin2) {
        if (L.k != 1) { set_error("fused second GEMM needs a 1x1 first convolution"); return IVOSW_ERR_INVALID; }
        fill_tap(P.taps[0], 0, 0, 1, L.cin);
        fill_tap(P.taps[1], 0, 0, L.stride2, L.cin2);
    } else {
        const int pad = L.dil * (L.k - 1) / 2;
        for (int kh = 0; kh < L.k; ++kh)
            for (int kw = 0; kw < L.k; ++kw)
                fill_tap(P.taps[kh * L.k + kw], kh * L.dil - pad, kw * L.dil - pad, L.stride, L.cin);
    }
    if (BN == 64) return launch_tc_variant<64, 4, true>(c, maps, P, s);
    if (pair) return launch_tc_variant<128, 4, true, true>(c, maps, P, s);
    return launch_tc_variant<128, 3, true>(c, maps, P, s);
}

// ------------------------------------------------------------------------------------------------
// layout helpers between the fp32 NHWC world (stem, probes) and the split-fp16 planes
// ------------------------------------------------------------------------------------------------
__global__ void split_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo,
                             long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= n) return;
    float2 v = *reinterpret_cast<const float2*>(in + i);
    v.x = fminf(fmaxf(v.x, -65504.f), 65504.f); v.y = fminf(fmaxf(v.y, -65504.f), 65504.f);
    __half2 h = __floats2half2_rn(v.x, v.y);
    float2 hf = __half22float2(h);
    *reinterpret_cast<__half2*>(hi + i) = h;
    *reinterpret_cast<__half2*>(lo + i) = __floats2half2_rn((v.x - hf.x) * 2048.0f, (v.y - hf.y) * 2048.0f);
}

__global__ void merge_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ out,
                             long long n, int use_lo) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= n) return;
    float2 h = __half22float2(*reinterpret_cast<const __half2*>(hi + i));
    float2 l = use_lo ? __half22float2(*reinterpret_cast<const __half2*>(lo + i)) : make_float2(0.f, 0.f);
    *reinterpret_cast<float2*>(out + i) = make_float2(fmaf(l.x, 1.0f / 2048.0f, h.x), fmaf(l.y, 1.0f / 2048.0f, h.y));
}

int launch_split(ivosw_ctx* c, const float* in, const SplitAct& out, long long n, cudaStream_t s) {
    split_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, s>>>(in, out.hi, out.lo, n);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

int launch_merge(ivosw_ctx* c, const SplitAct& in, float* out, long long n, int use_lo, cudaStream_t s) {
    merge_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, s>>>(in.hi, in.lo, out, n, use_lo);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    return IVOSW_OK;
}

}  // namespace ivosw
