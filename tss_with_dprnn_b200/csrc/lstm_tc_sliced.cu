// The fused tcgen05 LSTM layer with two half-jobs per CTA pair in ping-pong (lstm_tc_pp.cu) as a PERSISTENT kernel over
// TIME-SLICED jobs.
//
// Why: a layer is J = tiles x directions equally long pair-jobs (T steps each) and the step time does not depend on the
// rows per CTA, so with one job per CTA pair the layer takes ceil(J / 74) waves - 2 for the headline shape (J = 98 intra /
// 128 inter) where 1.32 / 1.73 would do: a quarter of the LSTM time is SMs waiting for the last wave.  Here every job is
// cut into k slices of ceil(T / k) steps; work items (job, slice) are handed out by an atomic ticket in slice-major order
// to <= 74 resident CTA pairs, so the layer takes ceil(k J / 74) / k waves (k = 3: 1.35 for J = 98; k = 4: 1.75 for 128).
//
// State across a slice boundary: the cell state c (fp32) goes through a small global scratch ([job][rank][unit][row],
// coalesced); h_{t-1} is re-read by TMA from the hb rows the previous slice stored.  A per-(job, rank) counter in global
// memory is released after the slice's stores have completed; the pair that draws (job, slice > 0) waits for it.  That
// predecessor always holds a SMALLER ticket, i.e. it is already running on some pair: no deadlock, whatever the residency
// of the grid.  Weights are reloaded only when the direction of the drawn item changes.
//
// The step itself is lstm_tc_pp_kernel's: half-job A = rows 0..63 of each CTA (epilogue warps 4..7), B = rows 64..127
// (warps 8..11), cta_group::2 MMAs with M = 128, the MMA thread alternating A(t), B(t), A(t+1), ...; results are bit for
// bit those of the one-job-per-pair kernels (tested).  All mbarrier phases are tracked with counters that run across
// items.  Inference only, uniform batches (the ragged inter-chunk layer and the training forward keep lstm_tc_pp_kernel).
#include "lstm_tc_common.cuh"

namespace dprnn {
using namespace tc;

constexpr uint32_t PS_BAR_BYTES = 256;
constexpr uint32_t PS_SM_TOTAL = SM_BAR + PS_BAR_BYTES;
static_assert(PS_SM_TOTAL <= 232448, "shared memory budget of one SM (227 KiB)");
constexpr uint32_t PS_HALF_ROWS = 64 * 128;   // byte offset of rows 64..127 inside a [128 x 128 B] tile

struct LstmSlicedParams {
    int T;                 // time steps per job
    int seq_dim;           // 2 (intra) or 1 (inter)
    int tiles_per_outer;
    int ndir;
    int njobs;             // tiles x outer x ndir
    int nslices;           // k
    int slice_len;         // ceil(T / k)
    int* ticket;           // [1]  zeroed by the launcher
    int* done;             // [njobs * 2] slices completed per (job, CTA rank); zeroed by the launcher
    float* cscratch;       // [njobs * 2][128 units][128 rows]
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

template <bool kFastAct, bool kF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
lstm_tc_sliced_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                      const __grid_constant__ CUtensorMap tmH64, const float* __restrict__ bias_perm,
                      const LstmSlicedParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* x_full = bars;                  // [NXS]  (leader's copy is the live one)
    uint64_t* x_empty = bars + NXS;           // [NXS]
    uint64_t* w_full = bars + 2 * NXS;
    uint64_t* d_full = bars + 2 * NXS + 1;    // [2 half-jobs][2 unit halves]
    uint64_t* h_free = bars + 2 * NXS + 5;    // [2]
    uint64_t* h_done = bars + 2 * NXS + 7;    // [2][2]  (leader's copy)
    uint64_t* hl_full = bars + 2 * NXS + 11;  // h_{t-1} tile of a continued job re-loaded from global
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NXS + 12);
    uint32_t* item_slot = tmem_slot + 1;
    float* sbias = reinterpret_cast<float*>(smem + SM_BIAS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int T = p.T;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) {
            printf("lstm_tc_sliced_kernel: dynamic shared memory base %u is not 1024-byte aligned\n", smem_u32(smem));
            __trap();
        }
        prefetch_tmap(&tmX); prefetch_tmap(&tmW); prefetch_tmap(&tmH64);
        for (int s = 0; s < NXS; ++s) { mbar_init(&x_full[s], 2); mbar_init(&x_empty[s], 1); }
        mbar_init(w_full, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&d_full[i], 1); mbar_init(&h_done[i], 8); }   // 4 warps x 2 CTAs
        mbar_init(&h_free[0], 1); mbar_init(&h_free[1], 1);
        mbar_init(hl_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<2>(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // counters that run across items (every role keeps the ones it needs; all advance identically)
    int cur_dir = -1;
    uint32_t n_wload = 0;       // completed weight loads                       (w_full phase)
    uint32_t n_hload = 0;       // completed h re-loads                         (hl_full phase)
    uint32_t x_it = 0;          // x ring slots produced / consumed             (x_full / x_empty phases)
    uint32_t g_step = 0;        // steps completed                              (d_full / h_free phases)
    uint32_t hd_ph = 0;         // h_done phases completed (one per step, plus one per continued item)
    const uint32_t total_items = (uint32_t)p.njobs * (uint32_t)p.nslices;

    for (;;) {
        // ---------------- draw a work item (slice-major order), the same for both CTAs of the pair
        if (threadIdx.x == 0 && rank == 0) {
            const uint32_t tk = (uint32_t)atomicAdd(p.ticket, 1);
            *item_slot = tk;
            st_shared_cluster_u32(map_to_cta(smem_u32(item_slot), 1), tk);
        }
        cluster_sync_all();                   // ticket visible in both CTAs; the previous item is finished in both
        const uint32_t item = *item_slot;
        if (item >= total_items) break;
        const int slice = (int)(item / (uint32_t)p.njobs), job = (int)(item % (uint32_t)p.njobs);
        const int dir = job % p.ndir, jt = job / p.ndir;
        const int outer = jt / p.tiles_per_outer;
        const int seq0 = (jt % p.tiles_per_outer) * 256 + (int)rank * 128;
        const int st0 = slice * p.slice_len;                                  // first step of the slice
        const int nsteps = min(p.slice_len, T - st0);                         // >= 1 by construction of nslices
        const uint32_t x_it0 = x_it;                                          // ring position at the start of the item
        auto t_of = [&](int g) { return dir ? T - 1 - g : g; };               // time index of global step g
        auto c1 = [&](int t, int sq) { return p.seq_dim == 2 ? t : sq; };
        auto c2 = [&](int t, int sq) { return p.seq_dim == 2 ? sq : t; };

        // ---------------- the previous slice of this job (a smaller ticket, hence running or done) must have stored its state;
        // then h_{t-1} = the rows it stored is re-read into this CTA's h tiles (same thread: acquire -> proxy fence -> TMA)
        if (slice > 0 && warp == 0 && elect_one()) {
            const int* flag = p.done + job * 2 + (int)rank;
            while (ld_acquire_gpu(flag) < slice) __nanosleep(64);
            asm volatile("fence.proxy.async;" ::: "memory");
            const int tp = t_of(st0 - 1);
            mbar_expect_tx(hl_full, 2 * TILE);
            for (int hj = 0; hj < 2; ++hj) {                                  // 64-row boxes (the store map)
                const int sq = seq0 + hj * 64;
                tma_load_4d(smem + SM_H + hj * PS_HALF_ROWS, &tmH64, hl_full, dir * 128, c1(tp, sq), c2(tp, sq), outer);
                tma_load_4d(smem + SM_H + TILE + hj * PS_HALF_ROWS, &tmH64, hl_full, dir * 128 + 64, c1(tp, sq), c2(tp, sq), outer);
            }
        }
        // ---------------- weights / biases of the item's direction
        const bool new_dir = dir != cur_dir;
        if (new_dir) {
            for (int i = threadIdx.x; i < 512; i += blockDim.x) sbias[i] = bias_perm[dir * 512 + i];
            if (warp == 0 && elect_one()) {
                mbar_expect_tx(w_full, 8 * TILE);
                const int wrow = ((dir * 2 + (int)rank) * 2) * 128;
                for (int nh = 0; nh < 2; ++nh)
                    for (int kb = 0; kb < 4; ++kb)
                        tma_load_2d(smem + SM_W + (nh * 4 + kb) * TILE, &tmW, w_full, kb * 64, wrow + nh * 128);
            }
            mbar_wait(w_full, n_wload & 1);
            ++n_wload;
            cur_dir = dir;
        }
        __syncthreads();                      // flag observed / biases written, for every warp of this CTA
        cluster_sync_all();                   // weights of both CTAs resident before the leader issues any MMA

        if (warp == 0) {
            // ================= TMA producer: x_t K-halves (128 rows: both half-jobs) into the ring =================
            if (elect_one()) {
                const uint32_t leader_full0 = map_to_cta(smem_u32(&x_full[0]), 0);
                uint32_t xi = x_it0;
                for (int ls = 0; ls < nsteps; ++ls) {
                    const int t = t_of(st0 + ls);
                    for (int half = 0; half < 2; ++half, ++xi) {
                        const uint32_t s = xi % NXS;
                        mbar_wait(&x_empty[s], ((xi / NXS) & 1) ^ 1);
                        const uint32_t lbar = leader_full0 + s * 8;
                        if (rank == 0) mbar_expect_tx_addr(smem_u32(&x_full[s]), 2 * TILE);
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
                        umma_bf16<2>(d, umma_desc_sw128(a_tile + j * PS_HALF_ROWS + kk * 32), umma_desc_sw128(b_tile + kk * 32),
                                     idesc, (first && kk == 0) ? 0u : 1u);
                };
                // h_done phase that publishes h_{t-1} for local step ls: a continued item starts with one pseudo-phase
                // (the re-loaded tiles); the phase of step ls-1 otherwise
                const uint32_t hd0 = hd_ph + (slice > 0 ? 1u : 0u);
                uint32_t xi = x_it0;
                for (int ls = 0; ls < nsteps; ++ls, xi += 2) {
                    const uint32_t s0 = xi % NXS, s1 = (xi + 1) % NXS;
                    mbar_wait_cluster(&x_full[s0], (xi / NXS) & 1);
                    mbar_wait_cluster(&x_full[s1], ((xi + 1) / NXS) & 1);
                    tc_fence_after();
                    const uint32_t x0 = aX + s0 * TILE, x1 = aX + s1 * TILE;
                    for (int j = 0; j < 2; ++j) {
                        if (ls == 0 && slice == 0) {                  // h_0 = 0: only the input projection
                            mma_kb(j, 0, 0, x0, true);  mma_kb(j, 0, 1, x1, false);
                            mma_kb(j, 1, 0, x0, true);  mma_kb(j, 1, 1, x1, false);
                            umma_commit_2cta(&h_free[j], 3);
                            umma_commit_2cta(&d_full[j * 2 + 0], 3); umma_commit_2cta(&d_full[j * 2 + 1], 3);
                            continue;
                        }
                        const uint32_t par = (hd0 + (uint32_t)ls - 1u) & 1u;
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
        } else if (warp >= 4) {
            // ================= epilogue: warps 4..7 = half-job A, 8..11 = half-job B =================
            const int e = warp - 4, j = e >> 2, q = e & 3;        // q = TMEM lane quadrant = warp % 4
            const int L = q * 32 + lane;                           // TMEM lane: rows 0..63 twice (2x2 layout)
            const int rih = L & 63, ub = L >> 6;                   // row inside the half-job, 32-unit block inside the unit half
            const int row = j * 64 + rih;                          // row inside the CTA's 128-sequence tile
            const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 2) * 128;
            const uint32_t leader_hdone = map_to_cta(smem_u32(&h_done[j * 2]), 0);
            const bool storer = ((e & 3) == 0 && lane == 0);
            const int bar_a = 1 + 2 * j, bar_b = 2 + 2 * j;
            const int sq = seq0 + j * 64;
            float c0[32], c1s[32];
            float* cs = p.cscratch + ((size_t)(job * 2 + (int)rank) * 128) * 128 + row;     // [unit][row]
            if (slice == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) { c0[i] = 0.f; c1s[i] = 0.f; }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    c0[i] = __ldcg(cs + (size_t)(ub * 32 + i) * 128);               // L2: the line may be stale in this SM's L1
                    c1s[i] = __ldcg(cs + (size_t)(64 + ub * 32 + i) * 128);
                }
                // the re-loaded h_{t-1} tiles stand in for the "h written" arrivals of a previous step
                mbar_wait(hl_full, n_hload & 1);
                __syncwarp();
                if (lane == 0) { mbar_arrive_remote(leader_hdone); mbar_arrive_remote(leader_hdone + 8); }
            }

            for (int ls = 0; ls < nsteps; ++ls) {
                const int t = t_of(st0 + ls);
                const uint32_t par = (g_step + (uint32_t)ls) & 1u;
                // ---------------- unit half 0
                mbar_wait(&d_full[j * 2 + 0], par);
                tc_fence_after();
                uint32_t pk[4][4];
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    lstm_cell8<kFastAct, false, 32, kF16>(tlane + 8 * g, sbias + ub * 32 + 8 * g, c0 + 8 * g, pk[g]);
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
                for (int g = 0; g < 4; ++g)
                    lstm_cell8<kFastAct, false, 32, kF16>(tlane + 128 + 8 * g, sbias + 256 + ub * 32 + 8 * g, c1s + 8 * g, pk[g]);
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
                    tma_store_4d(&tmH64, smem + SM_H + j * PS_HALF_ROWS, dir * 128, c1(t, sq), c2(t, sq), outer);
                    tma_store_4d(&tmH64, smem + SM_H + TILE + j * PS_HALF_ROWS, dir * 128 + 64, c1(t, sq), c2(t, sq), outer);
                    bulk_commit();
                }
            }
            if (slice + 1 < p.nslices) {      // hand the cell state to the next slice
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    cs[(size_t)(ub * 32 + i) * 128] = c0[i];
                    cs[(size_t)(64 + ub * 32 + i) * 128] = c1s[i];
                }
            }
            if (storer) {                     // the h rows of the slice are in global memory
                bulk_wait0();
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            __threadfence();                  // c scratch / h rows before the CTA barrier that precedes the release
        }
        // ---------------- every role advances the phase counters identically
        if (slice > 0) ++n_hload;
        g_step += (uint32_t)nsteps;
        hd_ph += (uint32_t)nsteps + (slice > 0 ? 1u : 0u);
        x_it = x_it0 + 2u * (uint32_t)nsteps;
        __syncwarp();
        tc_fence_before();
        __syncthreads();                      // stores of c and (storers) completion of the h stores precede the release
        if (threadIdx.x == 0 && slice + 1 < p.nslices) {
            __threadfence();
            st_release_gpu(p.done + job * 2 + (int)rank, slice + 1);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc<2>(tmem, 512);
}

static long sliced_njobs(int B, int S, int K, int inter, int ndir) {
    const long tiles = inter ? (long)((K + 255) / 256) * B : ((long)B * S + 255) / 256;
    return tiles * ndir;
}

}  // namespace dprnn

using namespace dprnn;

// The number of slices that minimises the layer's length in step-times over k <= kmax: the slices of a job are a chain, so
// the layer cannot be shorter than k * ceil(T / k) >= T, nor than the ceil(k J / pairs) rounds of ceil(T / k) steps the
// pairs need (ties: the smaller k - every slice boundary costs a state hand-off).
extern "C" int dprnn_lstm_sliced_auto(int B, int S, int K, int inter, int ndir, int pairs, int kmax) {
    const long J = sliced_njobs(B, S, K, inter, ndir);
    const int T = inter ? S : K;
    if (pairs <= 0) pairs = 74;
    long best = -1;
    int best_k = 1;
    for (int k = 1; k <= kmax && k <= T; ++k) {
        const long len = (T + k - 1) / k;
        const long ks = (T + len - 1) / len;                  // slices that actually hold a step
        const long rounds = (ks * J + pairs - 1) / pairs;
        const long cost = (rounds > ks ? rounds : ks) * len;
        if (best < 0 || cost < best) { best = cost; best_k = k; }
    }
    return best_k;
}

// Scheduler workspace of dprnn_lstm_layer_bf16_sliced: ticket + completion counters + the cell-state hand-off scratch.
extern "C" size_t dprnn_lstm_sliced_workspace_bytes(int B, int S, int K, int inter, int ndir) {
    const long njobs = sliced_njobs(B, S, K, inter, ndir);
    return 256 + (size_t)njobs * 2 * sizeof(int) + 256 + (size_t)njobs * 2 * 128 * 128 * sizeof(float);
}

// dprnn_lstm_layer_bf16_pp (same arguments, same weight packing, same results bit for bit) with persistent CTA pairs over
// k = nslices time slices per job (see the header of this file); nslices <= 0: chosen by dprnn_lstm_sliced_auto (k <= 8).
extern "C" int dprnn_lstm_layer_bf16_sliced(const void* x, const void* w_packed, const float* bias_perm, void* hout, int B,
                                            int S, int K, int inter, int hidden, int ndir, int flags, int nslices,
                                            int max_pairs, void* workspace, void* stream) {
    DPRNN_CHECK_ARG(x && w_packed && bias_perm && hout && workspace && B > 0 && S > 0 && K > 0);
    DPRNN_CHECK_ARG(hidden == 128 && (ndir == 1 || ndir == 2));
    DPRNN_CHECK_ARG(((uintptr_t)x | (uintptr_t)w_packed | (uintptr_t)hout | (uintptr_t)workspace) % 16 == 0);
    const int fast_act = flags & DPRNN_LSTM_FAST_ACT, f16 = flags & DPRNN_LSTM_FP16;
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tmX, tmW, tmH;
    const uint64_t ldx = 128 * 2, ldh = (uint64_t)ndir * 128 * 2;
    LstmSlicedParams p;
    p.ndir = ndir;
    uint64_t dX[4], sX[4], dH[4], sH[4];
    uint32_t box[4] = {64, 1, 1, 1}, boxh[4] = {64, 1, 1, 1};
    if (!inter) {      // [feat, t=k (K), seq=(b,s) (B*S), 1]
        dX[0] = 128; dX[1] = K; dX[2] = (uint64_t)B * S; dX[3] = 1;
        sX[0] = 2; sX[1] = ldx; sX[2] = (uint64_t)K * ldx; sX[3] = (uint64_t)B * S * K * ldx;
        sH[0] = 2; sH[1] = ldh; sH[2] = (uint64_t)K * ldh; sH[3] = (uint64_t)B * S * K * ldh;
        box[2] = 128; boxh[2] = 64;
        p.T = K; p.seq_dim = 2; p.tiles_per_outer = (int)(((long)B * S + 255) / 256);
    } else {           // [feat, seq=k (K), t=s (S), b (B)]
        dX[0] = 128; dX[1] = K; dX[2] = S; dX[3] = B;
        sX[0] = 2; sX[1] = ldx; sX[2] = (uint64_t)K * ldx; sX[3] = (uint64_t)S * K * ldx;
        sH[0] = 2; sH[1] = ldh; sH[2] = (uint64_t)K * ldh; sH[3] = (uint64_t)S * K * ldh;
        box[1] = 128; boxh[1] = 64;
        p.T = S; p.seq_dim = 1; p.tiles_per_outer = (K + 255) / 256;
    }
    p.njobs = (int)sliced_njobs(B, S, K, inter, ndir);
    for (int i = 0; i < 4; ++i) dH[i] = dX[i];
    dH[0] = (uint64_t)ndir * 128;
    const CUtensorMapDataType t16 = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (make_tmap(&tmX, t16, 4, x, dX, sX, box)) return 1;
    if (make_tmap(&tmH, t16, 4, hout, dH, sH, boxh)) return 1;
    const uint64_t dW[2] = {256, (uint64_t)ndir * 512}, sW[2] = {2, 512};
    const uint32_t bW[2] = {64, 128};
    if (make_tmap(&tmW, t16, 2, w_packed, dW, sW, bW)) return 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int pairs_cap = sms / 2;
    if (max_pairs > 0 && pairs_cap > max_pairs) pairs_cap = max_pairs;       // leave SMs to the kernels of other streams
    if (nslices <= 0) nslices = dprnn_lstm_sliced_auto(B, S, K, inter, ndir, pairs_cap, 8);
    // slices: every slice must hold at least one step
    if (nslices > p.T) nslices = p.T;
    p.slice_len = (p.T + nslices - 1) / nslices;
    p.nslices = (p.T + p.slice_len - 1) / p.slice_len;
    uint8_t* ws = (uint8_t*)workspace;
    p.ticket = (int*)ws;
    p.done = (int*)(ws + 256);
    const size_t done_bytes = ((size_t)p.njobs * 2 * sizeof(int) + 255) / 256 * 256;
    p.cscratch = (float*)(ws + 256 + done_bytes);
    DPRNN_CUDA(cudaMemsetAsync(ws, 0, 256 + done_bytes, st));
    long pairs = (long)p.njobs * p.nslices;
    if (pairs > pairs_cap) pairs = pairs_cap;
    auto kern = f16 ? (fast_act ? lstm_tc_sliced_kernel<true, true> : lstm_tc_sliced_kernel<false, true>)
                    : (fast_act ? lstm_tc_sliced_kernel<true, false> : lstm_tc_sliced_kernel<false, false>);
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_SM_TOTAL));
    kern<<<(unsigned)(pairs * 2), 384, PS_SM_TOTAL, st>>>(tmX, tmW, tmH, bias_perm, p);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
