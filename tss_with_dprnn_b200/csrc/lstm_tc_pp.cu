// The fused tcgen05 LSTM layer (lstm_tc.cu) with TWO HALF-JOBS per CTA pair in ping-pong.
//
// In lstm_tc_kernel the epilogue phases of a step run at the MUFU roofline (2.6 us of the 3.9 us step for 128 rows per
// CTA) and the rest is the hand-off h complete -> MMA -> commit -> TMEM load, during which the MUFU pipe idles.  A second
// independent job would fill that gap but does not fit the shared memory next to the resident weights - two HALF jobs
// do: the pair's 256 sequences are split into half-job A (rows 0..63 of each CTA) and B (rows 64..127).  Each half-job
// is a cta_group::2 MMA with M = 128 (64 rows per CTA); its accumulator for unit-half nh occupies 128 TMEM columns with
// the M = 128 "2x2" layout: lanes 0..63 hold N-columns 0..127, lanes 64..127 hold N-columns 128..255 of the same 64 rows.
// The weight rows are packed so that N-column c of MMA nh is  gate (c % 128) / 32  of unit  64 nh + 32 (c / 128) + c % 32:
// every lane then owns all four gates of 32 units.  Shared memory (weights, h and x tiles) and TMEM (4 x 128 columns) are
// exactly those of the one-job kernel; the x tiles (128 rows) serve both half-jobs.
// Warps 4..7 run half-job A's cell updates, warps 8..11 half-job B's, each with its own barriers; the MMA thread
// alternates A(t), B(t), A(t+1), ...: while it waits for A's h_t, B's MMAs are in flight and B's epilogue keeps the MUFU
// pipe busy, and vice versa.  Inference; uniform batches and the ragged inter-chunk layer (one pair-job per utterance).
//
// half_tiles (small batches): when the layer has so few sequences that 256-sequence tiles would leave more than half of the
// CTA pairs without a job, the tiles shrink to 128 sequences and every pair runs half-job A only - twice as many SMs work,
// and a step without a ping-pong partner is shorter than a step with one (a half-job's MMAs and cell update instead of two
// interleaved): cfg 4 (16 utterances per GPU), cfg 1 (B = 1) and the training forward at B = 16.
//
// kFuse (uniform batches, inference): the norm + residual that ends the PREVIOUS half-block (dprnn.py:90-92 / 98-99) is
// applied here, while the layer's input is loaded, instead of by a pass of its own over the residual stream (2.4 GB and
// 0.44 ms per half-block at B = 64, 13 % of the step, on a kernel that leaves 70 % of the HBM bandwidth unused).  Each CTA's
// TMA lands the OLD residual tile in the x ring; warps 2 and 3 - idle in the plain kernel - read the matching rows of y
// (the Linear output, straight from global memory, issued before the tile arrives), rewrite the tile in place as
// x + norm(y) with the arithmetic of norm_residual_bf16res_kernel (bit-identical), publish it to the tensor core, and the
// direction-0 job also stores it as the new residual stream (a second buffer: the other direction still reads the old one).
#include "lstm_tc_common.cuh"

namespace dprnn {
using namespace tc;

constexpr uint32_t PP_BAR_BYTES = 384;
constexpr uint32_t PP_SM_TOTAL = SM_BAR + PP_BAR_BYTES;
static_assert(PP_SM_TOTAL <= 232448, "shared memory budget of one SM (227 KiB)");
constexpr uint32_t HALF_ROWS = 64 * 128;      // byte offset of rows 64..127 inside a [128 x 128 B] tile

// The three tensor maps of the staged training epilogue (kStage): gates / cell state / fp32 h with the geometry of tmH64.
struct LstmSaveMaps { CUtensorMap g, c, h; };

template <bool kFastAct, bool kTrain, bool kF16, bool kFuse = false, bool kStage = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
lstm_tc_pp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmH64, const float* __restrict__ bias_perm, const LstmTcParams p,
                  const __grid_constant__ LstmSaveMaps sm) {
    static_assert(!kStage || kTrain, "the staged epilogue belongs to the training forward");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* x_full = bars;                  // [NXS]  (leader's copy is the live one)
    uint64_t* x_empty = bars + NXS;           // [NXS]
    uint64_t* w_full = bars + 2 * NXS;
    uint64_t* d_full = bars + 2 * NXS + 1;    // [2 half-jobs][2 unit halves]
    uint64_t* h_free = bars + 2 * NXS + 5;    // [2]
    uint64_t* h_done = bars + 2 * NXS + 7;    // [2][2]  (leader's copy)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NXS + 11);
    uint64_t* x_raw = bars + 2 * NXS + 12;    // [NXS]  kFuse: the old residual tile has landed (this CTA's own copy)
    uint64_t* sv_full = bars + 3 * NXS + 12;  // [4 warps][2 buffers]  kStage: a call's saved values are in the staging buffer
    uint64_t* sv_free = sv_full + 8;          // [4][2]  ... and its bulk stores have read them
    float* sbias = reinterpret_cast<float*>(smem + SM_BIAS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int job = blockIdx.x >> 1;
    const int dir = job % p.ndir;
    const int jt = job / p.ndir;
    int outer = jt / p.tiles_per_outer;
    const int nj = p.half_tiles ? 1 : 2;                       // half-jobs this pair runs
    const int seq0 = (jt % p.tiles_per_outer) * (128 * nj) + (int)rank * (64 * nj);
    const uint32_t xbytes = TILE / 2 * nj;                     // one K-half of this CTA's rows
    int T = p.T, t_base = 0;
    if (p.jobs) {          // ragged inter-chunk layer: one pair-job per utterance (its chunks start at t_base)
        const int2 jb = p.jobs[jt];
        t_base = jb.x; T = jb.y; outer = 0;
    }

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) {
            printf("lstm_tc_pp_kernel: dynamic shared memory base %u is not 1024-byte aligned\n", smem_u32(smem));
            __trap();
        }
        prefetch_tmap(&tmX); prefetch_tmap(&tmW); prefetch_tmap(&tmH64);
        if constexpr (kStage) { prefetch_tmap(&sm.g); prefetch_tmap(&sm.c); prefetch_tmap(&sm.h); }
        // x_full: plain = the two CTAs' TMA halves; kFuse = the two converter warps of both CTAs
        for (int s = 0; s < NXS; ++s) { mbar_init(&x_full[s], kFuse ? 4 : 2); mbar_init(&x_empty[s], 1); mbar_init(&x_raw[s], 1); }
        mbar_init(w_full, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&d_full[i], 1); mbar_init(&h_done[i], 8); }   // 4 warps x 2 CTAs
        mbar_init(&h_free[0], 1); mbar_init(&h_free[1], 1);
        if constexpr (kStage) for (int i = 0; i < 16; ++i) mbar_init(&sv_full[i], 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sbias[i] = bias_perm[dir * 512 + i];
    if (warp == 2) tmem_alloc<2>(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();

    if (warp == 0 && elect_one()) {
        mbar_expect_tx(w_full, 8 * TILE);
        const int wrow = ((dir * 2 + (int)rank) * 2) * 128;
        for (int nh = 0; nh < 2; ++nh)
            for (int kb = 0; kb < 4; ++kb)
                tma_load_2d(smem + SM_W + (nh * 4 + kb) * TILE, &tmW, w_full, kb * 64, wrow + nh * 128);
    }
    mbar_wait(w_full, 0);
    cluster_sync_all();
    const uint32_t tmem = *tmem_slot;

    auto c1 = [&](int t, int sq) { return p.seq_dim == 2 ? t : sq; };
    auto c2 = [&](int t, int sq) { return p.seq_dim == 2 ? sq : t_base + t; };

    if (warp == 0) {
        // ================= TMA producer: x_t K-halves (128 rows: both half-jobs) into the ring =================
        if (elect_one()) {
            const uint32_t leader_full0 = map_to_cta(smem_u32(&x_full[0]), 0);
            int it = 0;
            for (int step = 0; step < T; ++step) {
                const int t = dir ? T - 1 - step : step;
                for (int half = 0; half < 2; ++half, ++it) {
                    const int s = it % NXS;
                    mbar_wait(&x_empty[s], ((it / NXS) & 1) ^ 1);
                    if constexpr (kFuse) {     // lands in this CTA only; the converter warps publish it to the leader
                        mbar_expect_tx(&x_raw[s], TILE);
                        tma_load_4d(smem + SM_X + s * TILE, &tmX, &x_raw[s], half * 64, c1(t, seq0), c2(t, seq0), outer);
                        continue;
                    }
                    const uint32_t lbar = leader_full0 + s * 8;
                    if (rank == 0) mbar_expect_tx_addr(smem_u32(&x_full[s]), 2 * xbytes);
                    else mbar_arrive_remote(lbar);
                    tma_load_4d_pair(smem + SM_X + s * TILE, &tmX, lbar, half * 64, c1(t, seq0), c2(t, seq0), outer);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only): A(t), B(t), A(t+1), ... =================
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_h16(128, 256, kF16);
            const uint32_t aW = smem_u32(smem + SM_W), aH = smem_u32(smem + SM_H), aX = smem_u32(smem + SM_X);
            auto mma_kb = [&](int j, int nh, int kb, uint32_t a_tile, bool first) {
                const uint32_t b_tile = aW + (nh * 4 + kb) * TILE;
                const uint32_t d = tmem + (uint32_t)(j * 2 + nh) * 128;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16<2>(d, umma_desc_sw128(a_tile + j * HALF_ROWS + kk * 32), umma_desc_sw128(b_tile + kk * 32),
                                 idesc, (first && kk == 0) ? 0u : 1u);
            };
            int it = 0;
            for (int step = 0; step < T; ++step, it += 2) {
                const int s0 = it % NXS, s1 = (it + 1) % NXS;
                mbar_wait_cluster(&x_full[s0], (it / NXS) & 1);
                mbar_wait_cluster(&x_full[s1], ((it + 1) / NXS) & 1);
                tc_fence_after();
                const uint32_t x0 = aX + s0 * TILE, x1 = aX + s1 * TILE;
                for (int j = 0; j < nj; ++j) {
                    if (step == 0) {                              // h_0 = 0: only the input projection
                        mma_kb(j, 0, 0, x0, true);  mma_kb(j, 0, 1, x1, false);
                        mma_kb(j, 1, 0, x0, true);  mma_kb(j, 1, 1, x1, false);
                        umma_commit_2cta(&h_free[j], 3);
                        umma_commit_2cta(&d_full[j * 2 + 0], 3); umma_commit_2cta(&d_full[j * 2 + 1], 3);
                        continue;
                    }
                    const uint32_t par = (step - 1) & 1;
                    mbar_wait_cluster(&h_done[j * 2 + 0], par);   // D[j][0] drained, units 0..63 of h_{t-1} written
                    tc_fence_after();
                    mma_kb(j, 0, 0, x0, true);  mma_kb(j, 0, 1, x1, false);
                    mma_kb(j, 0, 2, aH, false);
                    mbar_wait_cluster(&h_done[j * 2 + 1], par);   // D[j][1] drained, h_{t-1} complete
                    tc_fence_after();
                    mma_kb(j, 0, 3, aH + TILE, false);
                    umma_commit_2cta(&d_full[j * 2 + 0], 3);
                    mma_kb(j, 1, 2, aH, true);   mma_kb(j, 1, 3, aH + TILE, false);
                    umma_commit_2cta(&h_free[j], 3);
                    mma_kb(j, 1, 0, x0, false);  mma_kb(j, 1, 1, x1, false);
                    umma_commit_2cta(&d_full[j * 2 + 1], 3);
                }
                umma_commit_2cta(&x_empty[s0], 3); umma_commit_2cta(&x_empty[s1], 3);
            }
        }
    } else if (kFuse && (warp == 2 || warp == 3)) {
        // ================= converter: x tile <- x + norm(y) in place (kFuse) =================
        const int ci = (warp - 2) * 32 + lane;                 // 64 threads: 16-byte chunk c of rows r0 + 8 k, k < 16
        const int c = ci & 7, r0 = ci >> 3;
        const uint32_t leader_full0 = map_to_cta(smem_u32(&x_full[0]), 0);
        // per job, 32-bit: row of (sequence seq0 + r0 + 8 k, time t) in the [rows, 128] buffers = base0 + k * kstride + t * tstride
        const bool intra = p.seq_dim == 2;
        const int kstride = intra ? 8 * p.K : 8, tstride = intra ? 1 : p.K;
        const int base0 = intra ? (seq0 + r0) * p.K : outer * p.S * p.K + seq0 + r0;
        const long left = p.seq_limit - seq0 - r0;             // valid rows: k < kmax
        const int kmax = left <= 0 ? 0 : (int)((left + 7) / 8 < 16 ? (left + 7) / 8 : 16);
        // utterance of every row as an offset from the first one, 4 bits each (a 128-row tile spans <= 16 utterances for
        // S >= 9; shorter utterances take the slow path below)
        const int b0 = intra ? (seq0 + r0) / p.S : outer;
        uint64_t boff = 0;
        bool wide = false;
        if (intra) {
            for (int k = 0; k < 16; ++k) {
                const int d = (seq0 + r0 + 8 * k) / p.S - b0;
                wide |= d > 15;
                boff |= (uint64_t)(d & 15) << (4 * k);
            }
        }
        int it = 0;
        for (int step = 0; step < T; ++step) {
            const int t = dir ? T - 1 - step : step;
#pragma unroll
            for (int half = 0; half < 2; ++half, ++it) {
                const int s = it % NXS;
                const int rbase = base0 + t * tstride;
                uint4 yv[16];
#pragma unroll
                for (int k = 0; k < 16; ++k)                   // y rows first: in flight while the tile lands
                    if (k < kmax) yv[k] = __ldg(p.fy + (long)(rbase + k * kstride) * 16 + half * 8 + c);
                float g[8], be[8];                             // affine of the thread's 8 channels of this K-half (L1 hits)
                {
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.fgamma + half * 64 + c * 8)), g1 = __ldg(reinterpret_cast<const float4*>(p.fgamma + half * 64 + c * 8 + 4));
                    const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.fbeta + half * 64 + c * 8)), e1 = __ldg(reinterpret_cast<const float4*>(p.fbeta + half * 64 + c * 8 + 4));
                    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
                    be[0] = e0.x; be[1] = e0.y; be[2] = e0.z; be[3] = e0.w; be[4] = e1.x; be[5] = e1.y; be[6] = e1.z; be[7] = e1.w;
                }
                mbar_wait(&x_raw[s], (it / NXS) & 1);
                uint8_t* tile = smem + SM_X + s * TILE;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (k >= kmax) continue;                   // padding rows of the last tile stay zero
                    const int row = r0 + 8 * k;
                    const int b = wide ? (seq0 + row) / p.S : b0 + (int)((boff >> (4 * k)) & 15);
                    const float2 st = __ldg(p.fmr + b);
                    uint4* px = reinterpret_cast<uint4*>(tile + sw128_offset(row, c));
                    const uint4 xv = *px;
                    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, yw[4] = {yv[k].x, yv[k].y, yv[k].z, yv[k].w};
                    uint32_t ob[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 xf = unpack_h16x2<kF16>(xw[q]), yf = unpack_h16x2<kF16>(yw[q]);
                        ob[q] = pack_h16x2<kF16>(norm_res1(xf.x, yf.x, st.x, st.y, g[2 * q], be[2 * q]),
                                                 norm_res1(xf.y, yf.y, st.x, st.y, g[2 * q + 1], be[2 * q + 1]));
                    }
                    const uint4 o = make_uint4(ob[0], ob[1], ob[2], ob[3]);
                    *px = o;
                    if (dir == 0) p.fxout[(long)(rbase + k * kstride) * 16 + half * 8 + c] = o;
                }
                fence_async_smem();                            // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(leader_full0 + s * 8);
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: warps 4..7 = half-job A, 8..11 = half-job B =================
        const int e = warp - 4, j = e >> 2, q = e & 3;        // q = TMEM lane quadrant = warp % 4
        if (j < nj) {
        const int L = q * 32 + lane;                           // TMEM lane: rows 0..63 twice (2x2 layout)
        const int rih = L & 63, ub = L >> 6;                   // row inside the half-job, 32-unit block inside the unit half
        const int row = j * 64 + rih;                          // row inside the CTA's 128-sequence tile
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 2) * 128;
        const uint32_t leader_hdone = map_to_cta(smem_u32(&h_done[j * 2]), 0);
        float c0[32], c1s[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { c0[i] = 0.f; c1s[i] = 0.f; }
        const bool storer = ((e & 3) == 0 && lane == 0);
        const int bar_a = 1 + 2 * j, bar_b = 2 + 2 * j;
        const int sq = seq0 + j * 64;
        const long seq = (long)seq0 + row;
        const bool live = kTrain && !kStage && seq < p.seq_limit;
        // kStage (half tiles only: half-job B's halves of the h tiles and of the x ring stages are unused, and so are its
        // epilogue warps): warp q owns the second 8 KiB of one of those tiles as two 4 KiB staging buffers and hands every
        // call's boxes to warp q + 4, which issues the bulk stores - the cell-update warps, the step's critical path, only
        // write shared memory
        uint8_t* stg_base = smem + (q < 2 ? SM_H + q * TILE : SM_X + (q - 2) * TILE) + HALF_ROWS;
        int ncall = 0;
        auto stage_buf = [&]() {                          // the buffer of two calls ago has been read by its bulk stores
            const int n = ncall++;
            if (n >= 2) mbar_wait(&sv_free[q * 2 + (n & 1)], ((n >> 1) - 1) & 1);
            return stg_base + (n & 1) * 4096;
        };
        auto stage_done = [&]() {                         // after the call's writes: publish to the async proxy, hand over
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sv_full[q * 2 + ((ncall - 1) & 1)]);
        };
        for (int step = 0; step < T; ++step) {
            const int t = dir ? T - 1 - step : step;
            const uint32_t par = step & 1;
            uint32_t* gd = nullptr;               // training: gates (packed bf16), c, h (fp32) of this row's 32 + 32 units
            float *cd = nullptr, *hd = nullptr;
            if constexpr (kTrain) {
                if (live) {
                    const long lr = p.seq_dim == 2 ? seq * p.K + t : ((long)outer * p.S + t) * p.K + seq;
                    gd = p.gates + (lr * p.ndir + dir) * 256 + ub * 64;       // uint32 units: 16 per 8-unit chunk
                    cd = p.cst + (lr * p.ndir + dir) * 128 + ub * 32;
                    hd = p.hf ? p.hf + (lr * p.ndir + dir) * 128 + ub * 32 : nullptr;
                }
            }
            // ---------------- unit half 0
            mbar_wait(&d_full[j * 2 + 0], par);
            tc_fence_after();
            uint32_t pk[4][4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if constexpr (kStage) {
                    uint8_t* buf = stage_buf();
                    lstm_cell8<kFastAct, kTrain, 32, kF16, true>(tlane + 8 * g, sbias + ub * 32 + 8 * g, c0 + 8 * g, pk[g],
                                                                 nullptr, nullptr, p.hf, buf, lane);
                    stage_done();
                } else {
                    lstm_cell8<kFastAct, kTrain, 32, kF16>(tlane + 8 * g, sbias + ub * 32 + 8 * g, c0 + 8 * g, pk[g],
                                                           gd ? gd + 16 * g : nullptr, cd + 8 * g, hd ? hd + 8 * g : nullptr);
                }
            }
            tc_fence_before();
            mbar_wait(&h_free[j], par);                // the MMAs that read this half-job's h_{t-1} have completed
            if (storer) bulk_wait_read0();             // ... and so has last step's TMA store of its h rows
            named_bar(bar_a, 128);
            {
                uint8_t* sH = smem + SM_H;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<uint4*>(sH + sw128_offset(row, ub * 4 + g)) = make_uint4(pk[g][0], pk[g][1], pk[g][2], pk[g][3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_hdone);
            // ---------------- unit half 1
            mbar_wait(&d_full[j * 2 + 1], par);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if constexpr (kStage) {
                    uint8_t* buf = stage_buf();
                    lstm_cell8<kFastAct, kTrain, 32, kF16, true>(tlane + 128 + 8 * g, sbias + 256 + ub * 32 + 8 * g, c1s + 8 * g,
                                                                 pk[g], nullptr, nullptr, p.hf, buf, lane);
                    stage_done();
                } else {
                    lstm_cell8<kFastAct, kTrain, 32, kF16>(tlane + 128 + 8 * g, sbias + 256 + ub * 32 + 8 * g, c1s + 8 * g, pk[g],
                                                           gd ? gd + 128 + 16 * g : nullptr, cd + 64 + 8 * g, hd ? hd + 64 + 8 * g : nullptr);
                }
            }
            tc_fence_before();
            {
                uint8_t* sH = smem + SM_H + TILE;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<uint4*>(sH + sw128_offset(row, ub * 4 + g)) = make_uint4(pk[g][0], pk[g][1], pk[g][2], pk[g][3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_hdone + 8);
            named_bar(bar_b, 128);
            if (storer) {
                tma_store_4d(&tmH64, smem + SM_H + j * HALF_ROWS, dir * 128, c1(t, sq), c2(t, sq), outer);
                tma_store_4d(&tmH64, smem + SM_H + TILE + j * HALF_ROWS, dir * 128 + 64, c1(t, sq), c2(t, sq), outer);
                bulk_commit();
            }
        }
        if (storer) bulk_wait0();
        } else if constexpr (kStage) {
            // ================= kStage: warps 8..11 store what warps 4..7 staged (same order of calls) =================
            if (lane == 0) {
                const int sq_w = seq0 + (q & 1) * 32, ub = q >> 1;
                const uint8_t* stg_base = smem + (q < 2 ? SM_H + q * TILE : SM_X + (q - 2) * TILE) + HALF_ROWS;
                int n = 0;
                for (int step = 0; step < T; ++step) {
                    const int t = dir ? T - 1 - step : step;
                    const int k1 = c1(t, sq_w), k2 = c2(t, sq_w);
                    for (int nh = 0; nh < 2; ++nh)
                        for (int g = 0; g < 4; ++g, ++n) {
                            const int unit0 = nh * 64 + ub * 32 + 8 * g;
                            const uint8_t* buf = stg_base + (n & 1) * 4096;
                            mbar_wait(&sv_full[q * 2 + (n & 1)], (n >> 1) & 1);
                            tma_store_4d(&sm.g, buf, dir * 256 + unit0 * 2, k1, k2, outer);
                            tma_store_4d(&sm.c, buf + 2048, dir * 128 + unit0, k1, k2, outer);
                            if (p.hf) tma_store_4d(&sm.h, buf + 3072, dir * 128 + unit0, k1, k2, outer);
                            bulk_commit();
                            if (n >= 1) {                  // the stores of the call before have read their buffer
                                bulk_wait_read1();
                                mbar_arrive(&sv_free[q * 2 + ((n - 1) & 1)]);
                            }
                        }
                }
                bulk_wait0();
            }
        }
    }
    __syncwarp();

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc<2>(tmem, 512);
}

}  // namespace dprnn

using namespace dprnn;

// Same arguments and results as dprnn_lstm_layer_bf16, except the weight packing: w_packed rows for direction d, CTA
// rank r and MMA nh are {[W_ih | W_hh][gate*H + 64*nh + 32*r + u] : gate = 0..3, u < 32} (Engine._pack_lstm_tc(half_jobs=True)).
static int lstm_pp_impl(const void* x, const void* w_packed, const float* bias_perm, void* hout, int B, int S, int K,
                        int inter, int hidden, int ndir, int flags, const int2* jobs, int n_jobs, void* stream,
                        void* gates = nullptr, float* cstate = nullptr, float* hout_f32 = nullptr,
                        const void* fy = nullptr, const float* fmr = nullptr, const float* fgamma = nullptr,
                        const float* fbeta = nullptr, void* fxout = nullptr) {
    const int fast_act = flags & DPRNN_LSTM_FAST_ACT, f16 = flags & DPRNN_LSTM_FP16;
    DPRNN_CHECK_ARG(!(f16 && gates));       // the training forward is built for bf16 operands (cfg 5: "bf16 gate GEMMs")
    DPRNN_CHECK_ARG(x && w_packed && bias_perm && hout && B > 0 && S > 0 && K > 0);
    DPRNN_CHECK_ARG(hidden == 128 && (ndir == 1 || ndir == 2));
    DPRNN_CHECK_ARG(((uintptr_t)x | (uintptr_t)w_packed | (uintptr_t)hout) % 16 == 0);
    CUtensorMap tmX, tmW, tmH;
    const uint64_t ldx = 128 * 2, ldh = (uint64_t)ndir * 128 * 2;
    LstmTcParams p{};
    p.ndir = ndir;
    uint64_t dX[4], sX[4], dH[4], sH[4];
    uint32_t box[4] = {64, 1, 1, 1}, boxh[4] = {64, 1, 1, 1};
    long njobs;
    if (!inter) {
        dX[0] = 128; dX[1] = K; dX[2] = (uint64_t)B * S; dX[3] = 1;
        sX[0] = 2; sX[1] = ldx; sX[2] = (uint64_t)K * ldx; sX[3] = (uint64_t)B * S * K * ldx;
        sH[0] = 2; sH[1] = ldh; sH[2] = (uint64_t)K * ldh; sH[3] = (uint64_t)B * S * K * ldh;
        box[2] = 128; boxh[2] = 64;
        p.T = K; p.seq_dim = 2; p.tiles_per_outer = (int)(((long)B * S + 255) / 256);
        njobs = (long)p.tiles_per_outer * ndir;
    } else {
        dX[0] = 128; dX[1] = K; dX[2] = S; dX[3] = B;
        sX[0] = 2; sX[1] = ldx; sX[2] = (uint64_t)K * ldx; sX[3] = (uint64_t)S * K * ldx;
        sH[0] = 2; sH[1] = ldh; sH[2] = (uint64_t)K * ldh; sH[3] = (uint64_t)S * K * ldh;
        box[1] = 128; boxh[1] = 64;
        p.T = S; p.seq_dim = 1; p.tiles_per_outer = (K + 255) / 256;
        njobs = (long)p.tiles_per_outer * B * ndir;
        if (jobs) {        // ragged: B == 1, S = total chunks of the packed batch, one pair-job per utterance and direction
            DPRNN_CHECK_ARG(B == 1 && K <= 256 && n_jobs > 0);
            njobs = (long)n_jobs * ndir;
        }
    }
    // small batches: 128-sequence tiles (half-job A only) when that still fits the CTA pairs in one wave
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    {
        const long nseq = inter ? K : (long)B * S;
        const long tiles128 = (nseq + 127) / 128, jobs128 = (inter ? tiles128 * B : tiles128) * ndir;
        const bool want = (flags & DPRNN_LSTM_HALF_TILES) || (!(flags & DPRNN_LSTM_FULL_TILES) && jobs128 <= sms / 2);
        if (want && !jobs && !fy) {
            p.half_tiles = 1;
            p.tiles_per_outer = (int)tiles128;
            njobs = jobs128;
            (inter ? box[1] : box[2]) = 64;
        }
    }
    p.jobs = jobs;
    p.gates = (uint32_t*)gates; p.cst = cstate; p.hf = hout_f32;
    p.fy = (const uint4*)fy; p.fmr = (const float2*)fmr; p.fgamma = fgamma; p.fbeta = fbeta; p.fxout = (uint4*)fxout;
    DPRNN_CHECK_ARG(!fy || (!gates && !jobs && fmr && fgamma && fbeta && fxout && fxout != x && (long)B * S * K < (1L << 27)));
    p.K = K; p.S = S;
    p.seq_limit = inter ? K : (long)B * S;
    for (int i = 0; i < 4; ++i) dH[i] = dX[i];
    dH[0] = (uint64_t)ndir * 128;
    const CUtensorMapDataType t16 = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (make_tmap(&tmX, t16, 4, x, dX, sX, box)) return 1;
    if (make_tmap(&tmH, t16, 4, hout, dH, sH, boxh)) return 1;
    const uint64_t dW[2] = {256, (uint64_t)ndir * 512}, sW[2] = {2, 512};
    const uint32_t bW[2] = {64, 128};
    if (make_tmap(&tmW, t16, 2, w_packed, dW, sW, bW)) return 1;
    // training forward on 128-sequence tiles: the saved gates / c / h leave through staging buffers and TMA (kStage)
    LstmSaveMaps sm{};
    const bool stage = gates && p.half_tiles && !(flags & DPRNN_LSTM_DIRECT_SAVE);
    if (stage) {
        uint64_t dS[4], sS[4];
        uint32_t bS[4] = {16, 1, 1, 1};
        bS[inter ? 1 : 2] = 32;
        for (int i = 0; i < 4; ++i) dS[i] = dX[i];
        auto scale = [&](uint64_t row_bytes) { sS[0] = 4; for (int i = 1; i < 4; ++i) sS[i] = sX[i] / ldx * row_bytes; };
        dS[0] = (uint64_t)ndir * 256; scale((uint64_t)ndir * 1024);
        if (make_tmap(&sm.g, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, gates, dS, sS, bS, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
        dS[0] = (uint64_t)ndir * 128; scale((uint64_t)ndir * 512); bS[0] = 8;
        if (make_tmap(&sm.c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, cstate, dS, sS, bS, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
        sm.h = sm.c;
        if (hout_f32 && make_tmap(&sm.h, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, hout_f32, dS, sS, bS, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
    }
    auto kern = stage ? (fast_act ? lstm_tc_pp_kernel<true, true, false, false, true> : lstm_tc_pp_kernel<false, true, false, false, true>)
                : gates ? (fast_act ? lstm_tc_pp_kernel<true, true, false> : lstm_tc_pp_kernel<false, true, false>)
                : fy  ? (f16 ? (fast_act ? lstm_tc_pp_kernel<true, false, true, true> : lstm_tc_pp_kernel<false, false, true, true>)
                             : (fast_act ? lstm_tc_pp_kernel<true, false, false, true> : lstm_tc_pp_kernel<false, false, false, true>))
                : f16 ? (fast_act ? lstm_tc_pp_kernel<true, false, true> : lstm_tc_pp_kernel<false, false, true>)
                      : (fast_act ? lstm_tc_pp_kernel<true, false, false> : lstm_tc_pp_kernel<false, false, false>);
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PP_SM_TOTAL));
    DPRNN_CHECK_ARG(njobs * 2 < (1L << 31));
    kern<<<(unsigned)(njobs * 2), 384, PP_SM_TOTAL, (cudaStream_t)stream>>>(tmX, tmW, tmH, bias_perm, p, sm);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_lstm_layer_bf16_pp(const void* x, const void* w_packed, const float* bias_perm, void* hout, int B,
                                        int S, int K, int inter, int hidden, int ndir, int fast_act, void* stream) {
    return lstm_pp_impl(x, w_packed, bias_perm, hout, B, S, K, inter, hidden, ndir, fast_act, nullptr, 0, stream);
}

// dprnn_lstm_layer_bf16_pp on the input x_in + norm(y) (GroupNorm(1,128) / gLN with per-utterance mean_rstd [B,2] and the
// affine gamma / beta [128]): the previous half-block's norm + residual applied while the input is loaded; x_out (a
// buffer other than x_in) receives the updated residual stream.  Bit-identical to dprnn_norm_residual_h16res followed by
// dprnn_lstm_layer_bf16_pp.
extern "C" int dprnn_lstm_layer_bf16_pp_fused(const void* x_in, const void* y, const float* mean_rstd, const float* gamma,
                                              const float* beta, void* x_out, const void* w_packed, const float* bias_perm,
                                              void* hout, int B, int S, int K, int inter, int hidden, int ndir, int flags,
                                              void* stream) {
    DPRNN_CHECK_ARG(y && mean_rstd && gamma && beta && x_out);
    DPRNN_CHECK_ARG(((uintptr_t)y | (uintptr_t)x_out) % 16 == 0);
    return lstm_pp_impl(x_in, w_packed, bias_perm, hout, B, S, K, inter, hidden, ndir, flags, nullptr, 0, stream, nullptr,
                        nullptr, nullptr, y, mean_rstd, gamma, beta, x_out);
}

// dprnn_lstm_inter_bf16_ragged with the half-job kernel (same arguments; half-job weight packing).
extern "C" int dprnn_lstm_inter_bf16_ragged_pp(const void* x, const void* w_packed, const float* bias_perm, void* hout,
                                               long total_chunks, int K, const void* utt_jobs, int n_utt, int hidden,
                                               int ndir, int fast_act, void* stream) {
    DPRNN_CHECK_ARG(utt_jobs && n_utt > 0 && total_chunks > 0 && total_chunks < (1L << 31));
    return lstm_pp_impl(x, w_packed, bias_perm, hout, 1, (int)total_chunks, K, 1, hidden, ndir, fast_act,
                        (const int2*)utt_jobs, n_utt, stream);
}

// dprnn_lstm_layer_bf16_train with the half-job kernel (half-job weight packing; same saved-state layouts).
extern "C" int dprnn_lstm_layer_bf16_train_pp(const void* x, const void* w_packed, const float* bias_perm, void* hout_bf16,
                                              void* gates_packed, float* cstate, float* hout_f32, int B, int S, int K,
                                              int inter, int hidden, int ndir, int fast_act, void* stream) {
    DPRNN_CHECK_ARG(gates_packed && cstate);        // hout_f32 may be NULL: h is then kept as bf16 (hout_bf16) only
    DPRNN_CHECK_ARG(((uintptr_t)gates_packed | (uintptr_t)cstate | (uintptr_t)hout_f32) % 32 == 0);
    return lstm_pp_impl(x, w_packed, bias_perm, hout_bf16, B, S, K, inter, hidden, ndir, fast_act, nullptr, 0, stream,
                        gates_packed, cstate, hout_f32);
}
