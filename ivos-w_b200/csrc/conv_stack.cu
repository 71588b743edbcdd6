// The whole res2..res5 convolution stack (52 bottleneck convolutions) as ONE persistent launch.
//
// conv_tc.cu runs one layer per kernel: every layer ends with the epilogue of its last tile and starts with a cold
// pipeline, the 3.46-wave layers of res4/res5 leave the SMs idle for half a wave, and every activation makes a full
// round trip through HBM between two kernels (the res2/res3 1x1 layers sit on the HBM roofline for exactly that
// reason; profiles/r1_ncu_conv_stack_v11.md).  Here the 148 CTAs walk ONE global list of (layer, m-tile, n-tile) work
// items in a fixed order; a tile starts as soon as the tiles it reads have been stored — tracked per (layer, m-tile)
// with a counter in global memory — so there is no grid-wide barrier anywhere in the stack:
//
//   * layer boundaries cost nothing: CTAs roll from the last tiles of one layer into the first tiles of the next;
//   * the list is ordered group by group inside a stage (res2: a few units at a time through all ten layers, then the
//     next group ...), so that what a layer writes is still in the 126 MB L2 when the next layer reads it: the
//     intermediate activations of res2/res3 stop travelling through HBM;
//   * a small unit batch (an 8-GPU frame shard) no longer pays 52 launches with few tiles each.
//
// The tile code itself is the staged-epilogue variant of conv_tc.cu (TMA -> 3-slot smem ring -> tcgen05.mma into two
// TMEM accumulator stages -> 16 epilogue warps -> swizzled staging buffer -> TMA store), with the tile width (64 / 128
// output channels), the tap list, the tensor maps and the BatchNorm vectors read per tile from a layer table in
// global memory, plus two store warps (one lane each, alternating tiles) that own the output stores: they send the staged
// half tiles, wait for the writes to COMPLETE (cp.async.bulk.wait_group 0, not .read) and only then publish the tile.
//
// Cross-CTA protocol (who may read what, when):
//   producer of a tile    TMA stores (async proxy) -> cp.async.bulk.wait_group 0 -> fence.proxy.async ->
//                         red.release.gpu.add  done[layer][m_tile] += 1            (store thread)
//   consumer of a tile    ld.acquire.gpu done[..] >= n_tiles(layer) for every m-tile its TMA boxes touch (3x3: the
//                         rows above / below, stride 2: the 2x finer input rows, residual: the same m-tile) ->
//                         fence.proxy.async -> TMA loads                                   (TMA producer thread)
// Deadlock freedom: tile t belongs to CTA t mod gridDim.x, every CTA works through its tiles in increasing order,
// and a tile only ever waits for tiles with a SMALLER index; all CTAs are resident (grid <= #SMs, one CTA per SM).
// Every buffer is written exactly once per launch (one output buffer per layer), so there are no WAR hazards to order.
// All waits are bounded and trap.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace ivosw {

constexpr int CS_STAGES = 3;
constexpr int CS_A_BYTES = TC_BM * TC_BK * 2;                       // 16 KB: one 128 x 64 fp16 plane
constexpr int CS_STAGE_BYTES = 4 * CS_A_BYTES;                      // A_hi, A_lo, W_hi, W_lo (W up to 128 rows)
constexpr int CS_STG_OFF = CS_STAGES * CS_STAGE_BYTES;              // staging buffer (hi + lo half tile, 32 KB)
constexpr int CS_STG_BYTES = 2 * CS_A_BYTES;
constexpr int CS_BAR_OFF = CS_STG_OFF + CS_STG_BYTES;
constexpr int CS_SMEM = CS_BAR_OFF + 256 + 1024 /*align slack*/;
constexpr int CS_EPI_WARPS = 16;
constexpr int CS_STORE_WARPS = 2;                                   // alternate tiles: one lane's completion wait overlaps the other's stores
constexpr int CS_THREADS = 32 * (2 + CS_EPI_WARPS + CS_STORE_WARPS);   // producer, MMA, 16 epilogue, store
static_assert(CS_SMEM <= 232448, "shared memory budget (227 KB)");

struct alignas(64) CsLayer {
    CUtensorMap a_hi, a_lo, w_hi, w_lo, r_hi, r_lo, o_hi, o_lo;
    TcTap taps[9];
    const float* scale;
    const float* shift;
    int bn, tiles_n, num_kb, has_res;          // (one 16-byte load for the MMA warp)
    int num_taps, cin_blocks, relu, out_hw;
    int in_hw, k, stride, pad;
    int in_layer, res_layer;                   // producing layers, -1 = produced before this launch
    int in_need, res_need;                     // n-tiles that complete one m-tile of those layers
    int in_m_tiles, pad0;
};

struct CsSegment { int layer, m0, ntiles, tile0; };                 // tiles [tile0, tile0 + ntiles): m-tiles from m0, n fastest

struct CsParams {
    const CsLayer* layers;
    const CsSegment* segs;
    int num_segs, total_tiles;
    int* done;                                 // [layer][done_stride] completed n-tiles per m-tile
    int done_stride;
    int terms;
    unsigned long long* sat_count;
    unsigned long long* prof;                  // measurement (IVOSW_STACK_PROFILE): [layer] cycles between tile completions
                                               // per CTA, [64 + layer] cycles the producer spent waiting for dependencies
};

struct CsTile { int layer, mt, nt; };

// tiles are visited in increasing order by every role: a cursor into the segment list is enough
__device__ __forceinline__ CsTile cs_decode(const CsParams& P, int t, int& cur, int& tiles_n_out) {
    int4 s = __ldg(reinterpret_cast<const int4*>(P.segs + cur));
    while (t >= s.w + s.z) { ++cur; s = __ldg(reinterpret_cast<const int4*>(P.segs + cur)); }
    const int tn = __ldg(&P.layers[s.x].tiles_n);
    const int local = t - s.w;
    tiles_n_out = tn;
    return CsTile{s.x, s.y + local / tn, local % tn};
}

__device__ __forceinline__ void cs_wait_done(const int* ctr, int need) {
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= need) return;
        if ((it & 255u) == 255u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > WAIT_TIMEOUT_CYCLES) __trap();
        }
        __nanosleep(64);
    }
}

template <bool PROF>
__global__ void __launch_bounds__(CS_THREADS, 1) conv_stack_kernel(const __grid_constant__ CsParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CS_BAR_OFF);
    uint64_t* full_bar = bars;                          // [3] operand block landed
    uint64_t* empty_bar = bars + 3;                     // [3] slot free again
    uint64_t* tfull_bar = bars + 6;                     // [2] accumulator stage complete
    uint64_t* tempty_bar = bars + 8;                    // [2] accumulator stage drained (16 warps)
    uint64_t* res_full = bars + 10;                     // [2] residual block of the r-th residual tile (r & 1)
    uint64_t* stg_full = bars + 12;                     // [2] staging buffer written (16 warps); index = tile parity in this CTA
    uint64_t* stg_empty = bars + 14;                    // [2] staging buffer read out by the TMA store (store lane of that parity)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool x3 = P.terms == 3;
    constexpr uint32_t TMEM_COLS = 512;                 // 2 accumulator stages x (D0 | D1) x 128 columns

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < CS_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], CS_EPI_WARPS);
            mbar_init(&res_full[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&stg_full[i], CS_EPI_WARPS); mbar_init(&stg_empty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        // Lane 0 issues the TMA loads of tile i.  Lanes 1..31 meanwhile poll the dependency counters of tile i + 1, one
        // counter per lane, so that the acquire round trips to L2 (~700 cycles each, up to five per tile: a 3x3 tile reads
        // three producer m-tiles, a stride-2 tile four, plus the residual's) are paid once, in parallel, and a tile ahead
        // of their use instead of serially in front of every tile's first load.
        int stage = 0; uint32_t phase = 0;
        int cur = 0, r_local = 0;
        int cur_next = 0;
        // what lanes >= 1 do: wait until everything tile `t` reads has been stored (see the protocol above)
        auto wait_deps = [&](int t) {
            int tiles_n;
            const CsTile T = cs_decode(P, t, cur_next, tiles_n);
            const CsLayer* L = P.layers + T.layer;
            const int in_layer = __ldg(&L->in_layer), res_layer = __ldg(&L->res_layer);
            const long long w0 = (PROF && lane == 1) ? clock64() : 0;
            int n_in = 0, mlo = 0;
            if (in_layer >= 0) {
                const int out_hw = __ldg(&L->out_hw), in_hw = __ldg(&L->in_hw), k = __ldg(&L->k), stride = __ldg(&L->stride),
                          pad = __ldg(&L->pad);
                const int pix_per_img = out_hw * out_hw;
                const int m0 = T.mt * TC_BM;
                const int n_img = m0 / pix_per_img;
                const int h0 = (m0 - n_img * pix_per_img) / out_hw;
                const int Hb = min(TC_BM / out_hw, out_hw);                 // output rows per image in this tile
                const int Nb = max(1, TC_BM / pix_per_img);                 // images per tile (8x8: 2)
                const int r_lo = max(0, stride * h0 - pad);
                const int r_hi = min(in_hw - 1, stride * (h0 + Hb - 1) - pad + (k - 1));
                const long long in_img = (long long)in_hw * in_hw;
                mlo = (int)((n_img * in_img + (long long)r_lo * in_hw) / TC_BM);
                int mhi = (int)(((n_img + Nb - 1) * in_img + (long long)r_hi * in_hw + in_hw - 1) / TC_BM);
                mhi = min(mhi, __ldg(&L->in_m_tiles) - 1);
                n_in = mhi - mlo + 1;
                const int need = __ldg(&L->in_need);
                const int* ctr = P.done + (long long)in_layer * P.done_stride;
                for (int j = lane - 1; j < n_in; j += 30) cs_wait_done(ctr + mlo + j, need);      // lanes 1..30
            }
            if (lane == 31 && __ldg(&L->has_res) && res_layer >= 0)
                cs_wait_done(P.done + (long long)res_layer * P.done_stride + T.mt, __ldg(&L->res_need));
            if (PROF && lane == 1) atomicAdd(P.prof + 64 + T.layer, (unsigned long long)(clock64() - w0));
        };
        if (lane >= 1 && (int)blockIdx.x < P.total_tiles) wait_deps(blockIdx.x);
        __syncwarp();
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
            if (lane >= 1) {
                if (tile + (int)gridDim.x < P.total_tiles) wait_deps(tile + gridDim.x);
            } else {
                int tiles_n;
                const CsTile T = cs_decode(P, tile, cur, tiles_n);
                const CsLayer* L = P.layers + T.layer;
                const int bn = __ldg(&L->bn), num_kb = __ldg(&L->num_kb), has_res = __ldg(&L->has_res);
                const int cin_blocks = __ldg(&L->cin_blocks), out_hw = __ldg(&L->out_hw);
                const int pix_per_img = out_hw * out_hw;
                const int m0 = T.mt * TC_BM;
                const int n_img = m0 / pix_per_img;
                const int h0 = (m0 - n_img * pix_per_img) / out_hw;
                asm volatile("fence.proxy.async;" ::: "memory");     // acquire (generic proxy) before the TMA reads (async proxy)
                // ---- operand blocks
                const uint32_t tx_bytes = x3 ? (uint32_t)(2 * CS_A_BYTES + 2 * bn * 128) : (uint32_t)(CS_A_BYTES + bn * 128);
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / cin_blocks, cb = kb - tap * cin_blocks;
                    const int4 tp = __ldg(reinterpret_cast<const int4*>(&L->taps[tap]));     // c_add, w_add, p, h_add
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * CS_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], tx_bytes);
                    const int c0 = tp.x + cb * TC_BK;
                    tma_load_5d(st, &L->a_hi, &full_bar[stage], c0, tp.y, tp.z, h0 + tp.w, n_img);
                    tma_load_2d(st + 2 * CS_A_BYTES, &L->w_hi, &full_bar[stage], kb * TC_BK, T.nt * bn);
                    if (x3) {
                        tma_load_5d(st + CS_A_BYTES, &L->a_lo, &full_bar[stage], c0, tp.y, tp.z, h0 + tp.w, n_img);
                        tma_load_2d(st + 2 * CS_A_BYTES + bn * 128, &L->w_lo, &full_bar[stage], kb * TC_BK, T.nt * bn);
                    }
                    if (++stage == CS_STAGES) { stage = 0; phase ^= 1; }
                }
                if (has_res) {
                    // residual block: one ring slot, completion on res_full (the MMA issuer steps over the slot)
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * CS_STAGE_BYTES;
                    uint64_t* rb = &res_full[r_local & 1];
                    mbar_expect_tx(rb, (x3 ? 2u : 1u) * (uint32_t)(bn / 64) * TC_BM * 128);
                    for (int g = 0; g < bn / 64; ++g) {
                        tma_load_2d(st + g * 2 * TC_BM * 128, &L->r_hi, rb, T.nt * bn + g * 64, m0);
                        if (x3) tma_load_2d(st + g * 2 * TC_BM * 128 + TC_BM * 128, &L->r_lo, rb, T.nt * bn + g * 64, m0);
                    }
                    if (++stage == CS_STAGES) { stage = 0; phase ^= 1; }
                    ++r_local;
                }
            }
            __syncwarp();          // tile + gridDim.x may be loaded: its inputs are complete (and ordered before lane 0's loads)
        }
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        const uint32_t smem_base = smem_u32(smem);
        int stage = 0; uint32_t full_bits = 0;
        int acc = 0; uint32_t acc_phase = 0;
        int cur = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
            int tiles_n;
            const CsTile T = cs_decode(P, tile, cur, tiles_n);
            const int4 li = __ldg(reinterpret_cast<const int4*>(&P.layers[T.layer].bn));     // bn, tiles_n, num_kb, has_res
            const int bn = li.x, num_kb = li.z;
            const uint32_t idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            const uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * bn) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            const bool prof = PROF && lane == 0;
            long long c0 = prof ? clock64() : 0, w_full = 0;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            if (prof) { const long long c1 = clock64(); atomicAdd(P.prof + 448 + T.layer, (unsigned long long)(c1 - c0)); }
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)(acc * 256);
            const uint32_t d1 = d0 + (uint32_t)bn;
            for (int kb = 0; kb < num_kb; ++kb) {
                if (prof) c0 = clock64();
                mbar_wait(&full_bar[stage], (full_bits >> stage) & 1u);
                if (prof) w_full += clock64() - c0;
                full_bits ^= 1u << stage;
                tc_fence_after();
                const uint32_t st = smem_base + (uint32_t)(stage * CS_STAGE_BYTES);
                const uint64_t a_hi = make_sw128_desc(st), a_lo = make_sw128_desc(st + CS_A_BYTES);
                const uint64_t b_hi = make_sw128_desc(st + 2 * CS_A_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                        const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
                        if (x3 && k == TC_BK / 16 - 1) {
                            // last K step: the longer instruction last, like conv_tc.cu (same accumulation order:
                            // the two kernels stay bit-identical)
                            umma_f16(d1, a_lo + adv, b_hi + adv, idesc, 1u);
                            umma_f16(d0, a_hi + adv, b_hi + adv, idesc_2n, 1u);
                        } else if (x3) {
                            umma_f16(d0, a_hi + adv, b_hi + adv, idesc_2n, accum);   // [D0 | D1] += A_hi [W_hi ; W_lo]^T
                            umma_f16(d1, a_lo + adv, b_hi + adv, idesc, 1u);         //       D1  += A_lo  W_hi^T
                        } else {
                            umma_f16(d0, a_hi + adv, b_hi + adv, idesc, accum);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == CS_STAGES) stage = 0;
            }
            if (prof) atomicAdd(P.prof + 384 + T.layer, (unsigned long long)w_full);
            if (li.w) { if (++stage == CS_STAGES) stage = 0; }        // the residual block's ring position
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp < 2 + CS_EPI_WARPS) {
        // ======================================= epilogue =======================================
        const int e = warp - 2;
        const int quarter = warp & 3;           // TMEM lanes this warp may touch: 32 * quarter .. + 31
        const int csub = e >> 2;                // which 16-column slice of a 64-column pass
        const int row = quarter * 32 + lane;
        uint32_t soff[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) soff[q] = (uint32_t)row * 128u + (uint32_t)(((2 * csub + q) ^ (row & 7)) << 4);
        int acc = 0; uint32_t acc_phase = 0;
        int stage = 0, r_local = 0, cur = 0;
        // the staging buffer is handed to the store lane of the tile's parity (see "store lanes" below); before it is
        // rewritten, the previous pass must have been read out
        uint32_t passes[2] = {0, 0};            // passes handed to each store lane so far
        int prev_owner = -1, t_cta = 0;
        __half2 sat = __float2half2_rn(0.f);            // running max of |hi| (fp16 range guard)
        const uint32_t stg_hi = smem_u32(smem + CS_STG_OFF), stg_lo = stg_hi + TC_BM * 128;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
            int tiles_n;
            const CsTile T = cs_decode(P, tile, cur, tiles_n);
            const CsLayer* L = P.layers + T.layer;
            const int4 li = __ldg(reinterpret_cast<const int4*>(&L->bn));
            const int bn = li.x, num_kb = li.z;
            const bool has_res = li.w != 0;
            // ReLU and the fp16 range guard as ONE clamp: [0, 65504] or [-65504, 65504]
            const float lo_clamp = __ldg(&L->relu) != 0 ? 0.f : -65504.f;
            const float* scale = L->scale;
            const float* shift = L->shift;
            const int np = bn >> 6;                                          // passes of 64 columns
            const bool prof = PROF && threadIdx.x == 64;
            const long long e0 = prof ? clock64() : 0;
            long long e_res = 0, e_tf = 0, e_stg = 0, ec = 0, e_ld = 0, e_math = 0, e_sts = 0, e_fence = 0;
            // The residual tile stays in its ring slot and is read pass by pass, right where it is added (keeping all of
            // it in registers from the start of the tile cost 32 registers per thread and pushed the epilogue into
            // local-memory spills, which the ~28 KB of L1 left beside 227 KB of shared memory cannot hold).
            stage = (stage + num_kb) % CS_STAGES;                            // ring position of the residual block, if any
            uint32_t rs = 0;
            if (has_res) {
                mbar_wait(&res_full[r_local & 1], (uint32_t)(r_local >> 1) & 1u);
                rs = smem_u32(smem + stage * CS_STAGE_BYTES);
                if (prof) e_res = clock64() - e0;
            }
            if (prof) ec = clock64();
            mbar_wait(&tfull_bar[acc], acc_phase);
            if (prof) e_tf = clock64() - ec;
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (g < np) {
                    const int n = T.nt * bn + g * 64 + csub * 16;
                    const uint32_t t_d0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + g * 64 + csub * 16);
                    uint32_t r0[16], r1[16];
                    if (prof) ec = clock64();
                    tmem_ld16(t_d0, r0);
                    if (x3) tmem_ld16(t_d0 + bn, r1);
                    // BatchNorm vectors of this thread's 16 channels: issued before the TMEM wait, their latency hides behind it
                    float sc[16], sh[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a4 = __ldg(reinterpret_cast<const float4*>(scale + n + q * 4));
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(shift + n + q * 4));
                        sc[q * 4] = a4.x; sc[q * 4 + 1] = a4.y; sc[q * 4 + 2] = a4.z; sc[q * 4 + 3] = a4.w;
                        sh[q * 4] = b4.x; sh[q * 4 + 1] = b4.y; sh[q * 4 + 2] = b4.z; sh[q * 4 + 3] = b4.w;
                    }
                    tmem_ld_wait();
                    if (prof) { const long long c = clock64(); e_ld += c - ec; ec = c; }
                    if (g == np - 1) {                  // last TMEM read of the tile: hand the accumulator stage back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    float v[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        float a = __uint_as_float(r0[k]);
                        if (x3) a = fmaf(__uint_as_float(r1[k]), 1.0f / 2048.0f, a);
                        v[k] = fmaf(a, sc[k], sh[k]);
                    }
                    if (has_res) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint4 h4 = lds128(rs + g * 2 * TC_BM * 128 + soff[q]);
                            const uint4 l4 = x3 ? lds128(rs + g * 2 * TC_BM * 128 + TC_BM * 128 + soff[q]) : make_uint4(0, 0, 0, 0);
                            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[u]));
                                const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[u]));
                                v[q * 8 + u * 2 + 0] += fmaf(lf.x, 1.0f / 2048.0f, hf.x);
                                v[q * 8 + u * 2 + 1] += fmaf(lf.y, 1.0f / 2048.0f, hf.y);
                            }
                        }
                        // (the slot is handed back to the TMA producer by the store lane, once every warp has arrived on
                        //  stg_full for the tile's last pass: those arrivals follow each warp's last read of the block and
                        //  its fence.proxy.async — no 512-thread barrier in the middle of the arithmetic)
                    }
                    uint32_t oh[8], ol[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float a = fminf(fmaxf(v[u * 2], lo_clamp), 65504.f);
                        const float b = fminf(fmaxf(v[u * 2 + 1], lo_clamp), 65504.f);
                        const __half2 h = __floats2half2_rn(a, b);
                        const float2 hf = __half22float2(h);
                        oh[u] = *reinterpret_cast<const uint32_t*>(&h);
                        sat = __hmax2(sat, __habs2(*reinterpret_cast<const __half2*>(&oh[u])));
                        ol[u] = pack_half2((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
                    }
                    const int owner = t_cta & (CS_STORE_WARPS - 1);
                    if (prof) { const long long c = clock64(); e_math += c - ec; ec = c; }
                    if (prev_owner >= 0) mbar_wait(&stg_empty[prev_owner], (passes[prev_owner] - 1) & 1u);   // previous half tile has left
                    if (prof) { const long long c = clock64(); e_stg += c - ec; ec = c; }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        sts128(stg_hi + soff[q], make_uint4(oh[q * 4], oh[q * 4 + 1], oh[q * 4 + 2], oh[q * 4 + 3]));
                        sts128(stg_lo + soff[q], make_uint4(ol[q * 4], ol[q * 4 + 1], ol[q * 4 + 2], ol[q * 4 + 3]));
                    }
                    if (prof) { const long long c = clock64(); e_sts += c - ec; ec = c; }
                    fence_proxy_async();                            // generic-proxy smem writes -> visible to the TMA store
                    if (prof) { const long long c = clock64(); e_fence += c - ec; ec = c; }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&stg_full[owner]);
                    passes[owner] += 1;
                    prev_owner = owner;
                }
            }
            if (has_res) { stage = (stage + 1) % CS_STAGES; ++r_local; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            ++t_cta;
            if (prof) {
                atomicAdd(P.prof + 128 + T.layer, (unsigned long long)e_res);
                atomicAdd(P.prof + 192 + T.layer, (unsigned long long)e_tf);
                atomicAdd(P.prof + 256 + T.layer, (unsigned long long)e_stg);
                atomicAdd(P.prof + 320 + T.layer, (unsigned long long)(clock64() - e0));
                atomicAdd(P.prof + 512 + T.layer, (unsigned long long)e_ld);
                atomicAdd(P.prof + 576 + T.layer, (unsigned long long)e_math);
                atomicAdd(P.prof + 640 + T.layer, (unsigned long long)e_sts);
                atomicAdd(P.prof + 704 + T.layer, (unsigned long long)e_fence);
            }
        }
        if (sat_hit(sat)) atomicAdd(P.sat_count, 1ull);
    } else if (lane == 0) {
        // ====================================== store lanes ======================================
        // Two store warps (one lane each): the first sends the even tiles of this CTA, the second the odd ones.  A tile may only be published once its writes have
        // COMPLETED (cp.async.bulk.wait_group 0 — about a microsecond after the stores were issued); bulk groups belong to
        // the issuing thread, so with one store thread that wait would sit between every two tiles' stores.  With two, the
        // other lane sends the next tile meanwhile.  Each lane has its own full / empty barrier pair, so neither can fall
        // more than one phase behind on a barrier it waits on.
        const int me = warp - (2 + CS_EPI_WARPS);
        int cur = 0, t_cta = 0, stage = 0;
        uint32_t my_pass = 0;
        long long t_prev = PROF ? clock64() : 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++t_cta) {
            int tiles_n;
            const CsTile T = cs_decode(P, tile, cur, tiles_n);
            const CsLayer* L = P.layers + T.layer;
            const int4 li = __ldg(reinterpret_cast<const int4*>(&L->bn));      // bn, tiles_n, num_kb, has_res
            stage = (stage + li.z) % CS_STAGES;                                 // ring position of the residual block, if any
            const int res_stage = stage;
            if (li.w) stage = (stage + 1) % CS_STAGES;
            if ((t_cta & (CS_STORE_WARPS - 1)) != me) continue;
            const int bn = li.x;
            const int np = bn >> 6;
            for (int g = 0; g < np; ++g) {
                mbar_wait(&stg_full[me], my_pass & 1u);
                ++my_pass;
                if (li.w && g == np - 1) mbar_arrive(&empty_bar[res_stage]);    // every warp is past its last residual read
                tma_store_2d(&L->o_hi, smem + CS_STG_OFF, T.nt * bn + g * 64, T.mt * TC_BM);
                if (x3) tma_store_2d(&L->o_lo, smem + CS_STG_OFF + TC_BM * 128, T.nt * bn + g * 64, T.mt * TC_BM);
                bulk_commit();
                bulk_wait_read0();
                mbar_arrive(&stg_empty[me]);
            }
            // publish the tile: its writes must have COMPLETED (not merely been read out of shared memory)
            bulk_wait0();
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(P.done + (long long)T.layer * P.done_stride + T.mt) : "memory");
            if (PROF) {
                const long long now = clock64();
                atomicAdd(P.prof + T.layer, (unsigned long long)(now - t_prev));
                t_prev = now;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side: the plan (layer table with tensor maps, segment list) and the launch
// ------------------------------------------------------------------------------------------------
struct StackPlan {
    int B = 0, terms = 0;
    const void* in_hi = nullptr;
    const void* arena = nullptr;
    unsigned long long weights_version = 0;
    int total_tiles = 0, num_segs = 0, done_stride = 0;
    CsLayer* layers_dev = nullptr;
    CsSegment* segs_dev = nullptr;
    int* done_dev = nullptr;
    size_t done_bytes = 0;
    unsigned long long stamp = 0;
};

struct StackState {
    std::vector<StackPlan> plans;
    unsigned long long clock = 0;
    bool attr_set = false;
    unsigned long long* prof_dev = nullptr;
};

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// bytes of one layer's output for one unit (fp32-equivalent: hi + lo planes)
static size_t layer_out_bytes(const ConvLayer& L) { return (size_t)L.out_hw * L.out_hw * L.cout * 4; }

void conv_stack_release(ivosw_ctx* c) {
    StackState* S = static_cast<StackState*>(c->stack_state);
    if (!S) return;
    for (StackPlan& p : S->plans) {
        if (p.layers_dev) cudaFree(p.layers_dev);
        if (p.segs_dev) cudaFree(p.segs_dev);
        if (p.done_dev) cudaFree(p.done_dev);
    }
    delete S;
    c->stack_state = nullptr;
}

// Output planes of layer li for a chunk of `cap` units live at a fixed offset of the arena.
static SplitAct arena_view(ivosw_ctx* c, size_t off_bytes, size_t bytes) {
    char* base = static_cast<char*>(c->stack_arena.p) + off_bytes;
    return SplitAct{reinterpret_cast<__half*>(base), reinterpret_cast<__half*>(base + bytes / 2)};
}

static int build_plan(ivosw_ctx* c, StackPlan& plan, int B, int terms, const SplitAct& in, const std::vector<size_t>& offs,
                      int cap) {
    int rc;
    const int nL = (int)c->layers.size();
    std::vector<CsLayer> tbl(nL);
    std::vector<int> m_tiles(nL), src_in(nL), src_res(nL);
    // who produces what: block input x (stack input, then every conv3 output), t1, t2, ds
    int x_layer = -1, t1_layer = -1, t2_layer = -1, ds_layer = -1;
    int max_m_tiles = 0;
    for (int li = 0; li < nL; ++li) {
        const ConvLayer& Lh = c->layers[li];
        CsLayer& L = tbl[li];
        memset(&L, 0, sizeof L);
        const long long M = (long long)B * Lh.out_hw * Lh.out_hw;
        m_tiles[li] = (int)((M + TC_BM - 1) / TC_BM);
        max_m_tiles = std::max(max_m_tiles, m_tiles[li]);
        int in_l, res_l = -1;
        if (Lh.first_of_block) in_l = x_layer;
        else if (Lh.k == 3) in_l = t1_layer;
        else if (Lh.is_downsample) in_l = x_layer;
        else { in_l = t2_layer; res_l = Lh.residual == 2 ? ds_layer : x_layer; }
        src_in[li] = in_l; src_res[li] = res_l;
        // 128-column tiles unless that leaves most SMs without a tile (small unit batches): same rule as conv_tc.cu
        const long long wide = (long long)m_tiles[li] * (Lh.cout / 128);
        const int bn = (Lh.cout >= 128 && wide > c->sm_count / 2) ? 128 : 64;
        L.bn = bn; L.tiles_n = Lh.cout / bn;
        L.num_taps = Lh.k * Lh.k; L.cin_blocks = Lh.cin / TC_BK; L.num_kb = L.num_taps * L.cin_blocks;
        L.has_res = (!Lh.first_of_block && Lh.k == 1 && !Lh.is_downsample) ? 1 : 0;
        L.relu = Lh.relu ? 1 : 0; L.out_hw = Lh.out_hw; L.in_hw = Lh.in_hw; L.k = Lh.k; L.stride = Lh.stride; L.pad = Lh.pad;
        L.scale = Lh.scale; L.shift = Lh.shift;
        L.in_layer = in_l; L.res_layer = L.has_res ? res_l : -1;
        const SplitAct src = in_l < 0 ? in : arena_view(c, offs[in_l], layer_out_bytes(c->layers[in_l]) * cap);
        const SplitAct dst = arena_view(c, offs[li], layer_out_bytes(Lh) * cap);
        const int K = Lh.k * Lh.k * Lh.cin;
        if ((rc = encode_act_map(&L.a_hi, src.hi, B, Lh.in_hw, Lh.cin, Lh.stride, Lh.out_hw))) return rc;
        if ((rc = encode_act_map(&L.a_lo, src.lo, B, Lh.in_hw, Lh.cin, Lh.stride, Lh.out_hw))) return rc;
        if ((rc = encode_w_map(&L.w_hi, Lh.w_hi, K, Lh.cout, bn))) return rc;
        if ((rc = encode_w_map(&L.w_lo, Lh.w_lo, K, Lh.cout, bn))) return rc;
        if ((rc = encode_out_map(&L.o_hi, dst.hi, M, Lh.cout))) return rc;
        if ((rc = encode_out_map(&L.o_lo, dst.lo, M, Lh.cout))) return rc;
        if (L.has_res) {
            const SplitAct rs = res_l < 0 ? in : arena_view(c, offs[res_l], layer_out_bytes(c->layers[res_l]) * cap);
            if ((rc = encode_out_map(&L.r_hi, rs.hi, M, Lh.cout))) return rc;
            if ((rc = encode_out_map(&L.r_lo, rs.lo, M, Lh.cout))) return rc;
        }
        for (int kh = 0; kh < Lh.k; ++kh)
            for (int kw = 0; kw < Lh.k; ++kw) {
                TcTap& t = L.taps[kh * Lh.k + kw];
                const int oy = kh - Lh.pad, ox = kw - Lh.pad;
                if (Lh.stride == 1) { t.c_add = 0; t.w_add = ox; t.p = 0; t.h_add = oy; }
                else {
                    const int py = oy & 1, px = ox & 1;
                    t.p = py; t.h_add = (oy - py) / 2; t.c_add = px * Lh.cin; t.w_add = (ox - px) / 2;
                }
            }
        if (Lh.first_of_block) t1_layer = li;
        else if (Lh.k == 3) t2_layer = li;
        else if (Lh.is_downsample) ds_layer = li;
        else x_layer = li;
    }
    for (int li = 0; li < nL; ++li) {
        CsLayer& L = tbl[li];
        L.in_need = src_in[li] >= 0 ? tbl[src_in[li]].tiles_n : 0;
        L.in_m_tiles = src_in[li] >= 0 ? m_tiles[src_in[li]] : 0;
        L.res_need = (L.has_res && src_res[li] >= 0) ? tbl[src_res[li]].tiles_n : 0;
    }
    // ---- the work list: stage by stage, inside a stage group by group (units), inside a group layer by layer
    struct StageSpec { int l0, l1, group; };
    std::vector<StageSpec> stages;
    {
        int l0 = 0;
        int sidx = 0;
        const int dflt[4] = {8, 16, 0, 0};      // units per group: res2, res3 (0 = the whole batch)
        const char* names[4] = {"IVOSW_STACK_G2", "IVOSW_STACK_G3", "IVOSW_STACK_G4", "IVOSW_STACK_G5"};
        for (int li = 1; li <= nL; ++li) {
            // a stage ends where the next bottleneck opens with a downsample branch (li + 2 is that branch)
            const bool boundary = li == nL || (c->layers[li].first_of_block && li + 2 < nL && c->layers[li + 2].is_downsample);
            if (boundary) {
                int g = env_int(names[sidx], dflt[sidx]);
                if (g <= 0 || g > B) g = B;
                if (g & 1) g += 1;              // 8x8 tiles span two images: keep group boundaries on tile boundaries
                stages.push_back({l0, li, g});
                l0 = li; sidx = std::min(sidx + 1, 3);
            }
        }
    }
    std::vector<CsSegment> segs;
    int tile0 = 0;
    for (const StageSpec& st : stages)
        for (int u0 = 0; u0 < B; u0 += st.group) {
            const int u1 = std::min(B, u0 + st.group);
            for (int li = st.l0; li < st.l1; ++li) {
                const long long mpu = (long long)c->layers[li].out_hw * c->layers[li].out_hw;
                const int m0 = (int)(u0 * mpu / TC_BM);
                const int m1 = (int)((u1 * mpu + TC_BM - 1) / TC_BM);
                CsSegment s{li, m0, (m1 - m0) * tbl[li].tiles_n, tile0};
                tile0 += s.ntiles;
                segs.push_back(s);
            }
        }
    segs.push_back(CsSegment{0, 0, 0x3fffffff, tile0});          // sentinel: the cursor never runs off the list
    plan.total_tiles = tile0; plan.num_segs = (int)segs.size();
    plan.done_stride = max_m_tiles;
    if (plan.layers_dev) cudaFree(plan.layers_dev);
    if (plan.segs_dev) cudaFree(plan.segs_dev);
    if (plan.done_dev) cudaFree(plan.done_dev);
    plan.layers_dev = nullptr; plan.segs_dev = nullptr; plan.done_dev = nullptr;
    IVOSW_CUDA(cudaMalloc(&plan.layers_dev, sizeof(CsLayer) * nL));
    IVOSW_CUDA(cudaMalloc(&plan.segs_dev, sizeof(CsSegment) * segs.size()));
    plan.done_bytes = sizeof(int) * (size_t)nL * max_m_tiles;
    IVOSW_CUDA(cudaMalloc(&plan.done_dev, plan.done_bytes));
    IVOSW_CUDA(cudaMemcpy(plan.layers_dev, tbl.data(), sizeof(CsLayer) * nL, cudaMemcpyHostToDevice));
    IVOSW_CUDA(cudaMemcpy(plan.segs_dev, segs.data(), sizeof(CsSegment) * segs.size(), cudaMemcpyHostToDevice));
    plan.B = B; plan.terms = terms; plan.in_hi = in.hi;
    return IVOSW_OK;
}

// Runs res2..res5 for B units on `in` (the pooled stem output, split-fp16 planes); *out receives the planes of r5 and
// stage_out[0..3] (optional) those of r2..r5.
int launch_conv_stack(ivosw_ctx* c, const SplitAct& in, int B, int terms, cudaStream_t s, SplitAct* out, SplitAct* stage_out) {
    int rc;
    if (!c->stack_state) c->stack_state = new StackState();
    StackState* S = static_cast<StackState*>(c->stack_state);
    const int nL = (int)c->layers.size();
    const int cap = std::max(B, c->chunk_cap_seen);
    c->chunk_cap_seen = cap;
    // arena: one output buffer per layer (nothing is overwritten inside a launch)
    std::vector<size_t> offs(nL);
    size_t total = 0;
    for (int li = 0; li < nL; ++li) { offs[li] = total; total += (layer_out_bytes(c->layers[li]) * cap + 1023) & ~(size_t)1023; }
    // (8x8 tiles of an odd unit count claim one more image: slack at the end, as in conv_tc.cu)
    total += layer_out_bytes(c->layers[nL - 1]) + (size_t)16 * 16 * 1024 * 4;
    if (c->stack_arena.bytes < total || c->stack_arena_cap != cap) {
        if (c->capturing) { set_error("conv stack workspace must grow during graph capture"); return IVOSW_ERR_STATE; }
        if ((rc = ensure(c->stack_arena, total))) return rc;
        c->stack_arena_cap = cap;
        for (StackPlan& p : S->plans) p.B = 0;          // every cached plan points into the old arena
    }
    // a plan bakes in: the unit count, the input planes, the arena and the weight buffers (tensor maps)
    StackPlan* plan = nullptr;
    for (StackPlan& p : S->plans)
        if (p.B == B && p.terms == terms && p.in_hi == in.hi && p.arena == c->stack_arena.p &&
            p.weights_version == c->assess_version) { plan = &p; break; }
    if (!plan) {
        if (c->capturing) { set_error("conv stack plan must be built outside graph capture"); return IVOSW_ERR_STATE; }
        if (S->plans.size() < 12) { S->plans.emplace_back(); plan = &S->plans.back(); }
        else plan = &*std::min_element(S->plans.begin(), S->plans.end(),
                                       [](const StackPlan& a, const StackPlan& b) { return a.stamp < b.stamp; });
        if ((rc = build_plan(c, *plan, B, terms, in, offs, cap))) { plan->B = 0; return rc; }
        plan->arena = c->stack_arena.p; plan->weights_version = c->assess_version;
    }
    plan->stamp = ++S->clock;
    if (!S->attr_set) {
        IVOSW_CUDA(cudaFuncSetAttribute(conv_stack_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
        IVOSW_CUDA(cudaFuncSetAttribute(conv_stack_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
        S->attr_set = true;
    }
    IVOSW_CUDA(cudaMemsetAsync(plan->done_dev, 0, plan->done_bytes, s));
    CsParams P;
    P.layers = plan->layers_dev; P.segs = plan->segs_dev; P.num_segs = plan->num_segs; P.total_tiles = plan->total_tiles;
    P.done = plan->done_dev; P.done_stride = plan->done_stride; P.terms = terms; P.sat_count = c->sat_count;
    P.prof = nullptr;
    static const bool profile = env_int("IVOSW_STACK_PROFILE", 0) != 0;
    if (profile && !c->capturing) {
        if (!S->prof_dev) IVOSW_CUDA(cudaMalloc(&S->prof_dev, sizeof(unsigned long long) * 768));
        IVOSW_CUDA(cudaMemsetAsync(S->prof_dev, 0, sizeof(unsigned long long) * 768, s));
        P.prof = S->prof_dev;
    }
    const int grid = std::min(plan->total_tiles, c->sm_count);
    if (P.prof) conv_stack_kernel<true><<<grid, CS_THREADS, CS_SMEM, s>>>(P);
    else conv_stack_kernel<false><<<grid, CS_THREADS, CS_SMEM, s>>>(P);
    c->launches += 1;
    IVOSW_CUDA(cudaGetLastError());
    if (P.prof) {       // measurement only: synchronous read-back and a table on stderr
        unsigned long long h[768];
        IVOSW_CUDA(cudaMemcpyAsync(h, S->prof_dev, sizeof h, cudaMemcpyDeviceToHost, s));
        IVOSW_CUDA(cudaStreamSynchronize(s));
        unsigned long long tot = 0, totw = 0;
        for (int li = 0; li < nL; ++li) { tot += h[li]; totw += h[64 + li]; }
        (void)tot;
        fprintf(stderr, "conv_stack profile (B = %d, grid = %d): cycles per CTA, per layer\n"
                        "  layer                          epilogue-total  = wait-res + wait-acc + wait-staging + work | mma: wait-operands wait-acc-free | dep-wait(lookahead)\n", B, grid);
        double te = 0;
        for (int li = 0; li < nL; ++li) {
            const ConvLayer& Lh = c->layers[li];
            const double e = (double)h[320 + li] / grid, r = (double)h[128 + li] / grid, t = (double)h[192 + li] / grid,
                         g = (double)h[256 + li] / grid;
            te += e;
            fprintf(stderr, "  %2d %dx%d s%d %4d->%4d @%2d  %9.0f = %8.0f + %8.0f + %8.0f + %8.0f | %8.0f %8.0f | %8.0f || ldtm %7.0f math %7.0f sts %7.0f fence %7.0f\n", li, Lh.k, Lh.k,
                    Lh.stride, Lh.cin, Lh.cout, Lh.out_hw, e, r, t, g, e - r - t - g, (double)h[384 + li] / grid,
                    (double)h[448 + li] / grid, (double)h[64 + li] / grid, (double)h[512 + li] / grid, (double)h[576 + li] / grid,
                    (double)h[640 + li] / grid, (double)h[704 + li] / grid);
        }
        fprintf(stderr, "  total epilogue-thread cycles per CTA %.0f, dependency wait %.0f\n", te, (double)totw / grid);
    }
    if (out) *out = arena_view(c, offs[nL - 1], layer_out_bytes(c->layers[nL - 1]) * cap);
    if (stage_out) {
        int k = 0;
        for (int li = 0; li < nL; ++li) {
            const bool stage_end = (li + 1 == nL) || (li + 3 < nL && c->layers[li + 3].is_downsample && !c->layers[li].first_of_block &&
                                                      c->layers[li].k == 1 && !c->layers[li].is_downsample);
            if (stage_end && k < 4) stage_out[k++] = arena_view(c, offs[li], layer_out_bytes(c->layers[li]) * cap);
        }
    }
    return IVOSW_OK;
}

}  // namespace ivosw
