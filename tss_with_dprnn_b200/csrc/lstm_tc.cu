// Fused LSTM layer on 5th-gen tensor cores (bf16 mode): per time step ONE tcgen05 accumulation
//     gates[256 seq, 512] = [x_t | h_{t-1}] (bf16, K=256) @ [W_ih | W_hh]^T      (fp32 accumulators in TMEM)
// followed by the cell update in registers.  The gate pre-activations never touch HBM: the time-parallel
// input projection x_t W_ih^T is fed by TMA (one 128x128 bf16 tile of x per step and CTA) and issued as
// the first half of the K loop, the recurrent half h_{t-1} W_hh^T reads the h tile the epilogue warps
// wrote to shared memory one step earlier.
//
// A CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns 256 sequences of one direction for all T steps:
//   * each CTA holds its own 128 sequences (A operand rows, TMEM lanes, cell state) and HALF of the weight
//     rows (B operand): 256 gate columns x K=256 in bf16 = 128 KiB - which is why the pair is needed: the
//     full [W_ih | W_hh] (256 KiB) does not fit one SM, and the hardware shares the halves across the pair;
//   * gate columns are permuted so that MMA instruction nh (N=256) produces, for hidden units 64nh..64nh+63,
//     the four gates at TMEM columns nh*256 + {0,64,128,192} + j;
//   * warp roles: 0 = TMA producer, 1 = MMA issuer (leader CTA, one elected thread), 2 = TMEM allocator,
//     4..11 = epilogue (warp%4 = TMEM lane quadrant, (warp-4)/4 = unit half).  Cell state: 64 fp32 registers
//     per epilogue thread.  h_t goes to shared memory as the next step's A operand (128B-swizzled, bf16) and
//     from there to HBM with one TMA store per step.
#include "lstm_tc_common.cuh"

namespace dprnn {
using namespace tc;

template <bool kFastAct, bool kTrain>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * NEPI, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmH, const float* __restrict__ bias_perm, const LstmTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* x_full = bars;                  // [NXS]  (leader's copy is the live one)
    uint64_t* x_empty = bars + NXS;           // [NXS]
    uint64_t* w_full = bars + 2 * NXS;
    uint64_t* d_full = bars + 2 * NXS + 1;    // [2]  gates of unit-half nh are complete in TMEM
    uint64_t* h_free = bars + 2 * NXS + 3;    //      every MMA that reads h_{t-1} has completed
    uint64_t* h_done = bars + 2 * NXS + 4;    // [2]  (leader's copy) D[nh] drained and h half nh written, both CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NXS + 6);
    float* sbias = reinterpret_cast<float*>(smem + SM_BIAS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int job = blockIdx.x >> 1;
    const int dir = job % p.ndir;
    const int jt = job / p.ndir;
    int outer = jt / p.tiles_per_outer;
    const int seq0 = (jt % p.tiles_per_outer) * 256 + (int)rank * 128;
    int T = p.T, t_base = 0;
    if (p.jobs) {          // one pair-job per utterance: its S_b chunks start at chunk t_base of the packed chunk space
        const int2 j = p.jobs[jt];
        t_base = j.x; T = j.y; outer = 0;
    }

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) {         // SWIZZLE_128B tiles need a 1024-byte aligned base
            printf("lstm_tc_kernel: dynamic shared memory base %u is not 1024-byte aligned\n", smem_u32(smem));
            __trap();
        }
        prefetch_tmap(&tmX); prefetch_tmap(&tmW); prefetch_tmap(&tmH);
        for (int s = 0; s < NXS; ++s) { mbar_init(&x_full[s], 2); mbar_init(&x_empty[s], 1); }
        mbar_init(w_full, 1);
        mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1);
        mbar_init(h_free, 1);
        mbar_init(&h_done[0], 2 * NEPI); mbar_init(&h_done[1], 2 * NEPI);   // one elected lane per epilogue warp, both CTAs
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sbias[i] = bias_perm[dir * 512 + i];
    if (warp == 2) tmem_alloc<2>(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();                   // barriers of both CTAs initialised before any remote arrive / TMA signal
    tc_fence_after();

    if (warp == 0 && elect_one()) {       // weights: this CTA's half of the rows of every (nh, kb) tile
        mbar_expect_tx(w_full, 8 * TILE);
        const int wrow = ((dir * 2 + (int)rank) * 2) * 128;
        for (int nh = 0; nh < 2; ++nh)
            for (int kb = 0; kb < 4; ++kb)
                tma_load_2d(smem + SM_W + (nh * 4 + kb) * TILE, &tmW, w_full, kb * 64, wrow + nh * 128);
    }
    mbar_wait(w_full, 0);
    cluster_sync_all();                   // both halves of the weights are resident
    const uint32_t tmem = *tmem_slot;

    // coordinates of (this CTA's 128 sequences, time t): intra (f, t, seq, 0) / inter (f, seq, t, outer)
    auto c1 = [&](int t) { return p.seq_dim == 2 ? t : seq0; };
    auto c2 = [&](int t) { return p.seq_dim == 2 ? seq0 : t_base + t; };

    if (warp == 0) {
        // ================= TMA producer: x_t K-halves into the ring =================
        if (elect_one()) {
            const uint32_t leader_full0 = map_to_cta(smem_u32(&x_full[0]), 0);
            int it = 0;
            for (int step = 0; step < T; ++step) {
                const int t = dir ? T - 1 - step : step;
                for (int half = 0; half < 2; ++half, ++it) {
                    const int s = it % NXS;
                    mbar_wait(&x_empty[s], ((it / NXS) & 1) ^ 1);
                    const uint32_t lbar = leader_full0 + s * 8;
                    if (rank == 0) mbar_expect_tx_addr(smem_u32(&x_full[s]), 2 * TILE);
                    else mbar_arrive_remote(lbar);
                    tma_load_4d_pair(smem + SM_X + s * TILE, &tmX, lbar, half * 64, c1(t), c2(t), outer);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        // TMEM: D0 = columns [0,256) (hidden units 0..63, gates i|f|g|o), D1 = [256,512) (units 64..127).
        // Per step:  x-part(D0) and the first K-half of h-part(D0) (units 0..63 of h_{t-1}) are issued as soon as the
        // epilogue has drained D0 and written that half (during its work on D1); when h_{t-1} is complete:
        // second K-half of h-part(D0) -> d_full[0]; h-part(D1) -> h_free; x-part(D1) -> d_full[1].
        // The epilogue of D0 therefore overlaps the MMAs of D1, and the next x-part(D0) the epilogue of D1.
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
            const uint32_t aW = smem_u32(smem + SM_W), aH = smem_u32(smem + SM_H), aX = smem_u32(smem + SM_X);
            auto mma_kb = [&](int nh, int kb, uint32_t a_tile, bool first) {
                const uint32_t b_tile = aW + (nh * 4 + kb) * TILE;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16<2>(tmem + nh * 256, umma_desc_sw128(a_tile + kk * 32), umma_desc_sw128(b_tile + kk * 32),
                                 idesc, (first && kk == 0) ? 0u : 1u);
            };
            int it = 0;
            for (int step = 0; step < T; ++step, it += 2) {
                const int s0 = it % NXS, s1 = (it + 1) % NXS;
                mbar_wait_cluster(&x_full[s0], (it / NXS) & 1);
                mbar_wait_cluster(&x_full[s1], ((it + 1) / NXS) & 1);
                tc_fence_after();
                if (step == 0) {                                  // h_0 = 0: only the input projection
                    mma_kb(0, 0, aX + s0 * TILE, true);  mma_kb(0, 1, aX + s1 * TILE, false);
                    mma_kb(1, 0, aX + s0 * TILE, true);  mma_kb(1, 1, aX + s1 * TILE, false);
                    umma_commit_2cta(&x_empty[s0], 3); umma_commit_2cta(&x_empty[s1], 3);
                    umma_commit_2cta(h_free, 3);
                    umma_commit_2cta(&d_full[0], 3); umma_commit_2cta(&d_full[1], 3);
                    continue;
                }
                const uint32_t par = (step - 1) & 1;
                mbar_wait_cluster(&h_done[0], par);               // D0 drained by both CTAs
                tc_fence_after();
                mma_kb(0, 0, aX + s0 * TILE, true);  mma_kb(0, 1, aX + s1 * TILE, false);
                mma_kb(0, 2, aH, false);                          // units 0..63 of h_{t-1} were published with h_done[0]
                mbar_wait_cluster(&h_done[1], par);               // D1 drained, h_{t-1} complete in both CTAs
                tc_fence_after();
                mma_kb(0, 3, aH + TILE, false);                   // only K=64 of D0 is left on the step's critical path
                umma_commit_2cta(&d_full[0], 3);
                mma_kb(1, 2, aH, true);   mma_kb(1, 3, aH + TILE, false);
                umma_commit_2cta(h_free, 3);
                mma_kb(1, 0, aX + s0 * TILE, false);  mma_kb(1, 1, aX + s1 * TILE, false);
                umma_commit_2cta(&x_empty[s0], 3); umma_commit_2cta(&x_empty[s1], 3);
                umma_commit_2cta(&d_full[1], 3);
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: gates -> (c, h) =================
        // all NEPI warps work on D0, then on D1: warp -> TMEM lane quadrant q = warp % 4, UPT-unit sub-block sub = e / 4
        // (16 warps x 16 units per thread measured the same step time as 8 x 32: the step is bound by the MUFU pipe and
        // the MMA <-> epilogue hand-off, not by the per-thread chain)
        const int e = warp - 4, q = e & 3, sub = e >> 2;
        const int row = q * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + sub * UPT;
        const uint32_t leader_hdone = map_to_cta(smem_u32(&h_done[0]), 0);
        float c0[UPT], c1s[UPT];
#pragma unroll
        for (int i = 0; i < UPT; ++i) { c0[i] = 0.f; c1s[i] = 0.f; }
        const bool storer = (warp == 4 && lane == 0);
        const long seq = (long)seq0 + row;
        const bool live = kTrain && seq < p.seq_limit;

        for (int step = 0; step < T; ++step) {
            const int t = dir ? T - 1 - step : step;
            const uint32_t par = step & 1;
            uint32_t* gd = nullptr;
            float *cd = nullptr, *hd = nullptr;
            if constexpr (kTrain) {
                if (live) {
                    const long lr = p.seq_dim == 2 ? seq * p.K + t : ((long)outer * p.S + t) * p.K + seq;
                    gd = p.gates + (lr * p.ndir + dir) * 256 + sub * (2 * UPT);      // uint32 units: 16 per 8-unit chunk
                    cd = p.cst + (lr * p.ndir + dir) * 128 + sub * UPT;
                    hd = p.hf + (lr * p.ndir + dir) * 128 + sub * UPT;
                }
            }
            // ---------------- unit half 0
            mbar_wait(&d_full[0], par);
            tc_fence_after();
            uint32_t pk[UPT / 8][4];
#pragma unroll
            for (int g = 0; g < UPT / 8; ++g)
                lstm_cell8<kFastAct, kTrain>(tbase + 8 * g, sbias + sub * UPT + 8 * g, c0 + 8 * g, pk[g],
                                             gd ? gd + 16 * g : nullptr, cd + 8 * g, hd + 8 * g);
            tc_fence_before();                         // our tcgen05.ld of D0 are complete
            mbar_wait(h_free, par);                    // the MMAs that read h_{t-1} have completed
            if (storer) bulk_wait_read0();             // ... and so has last step's TMA store of the h tile
            named_bar(1, 32 * NEPI);
            {
                uint8_t* sH = smem + SM_H;             // K-block 0 = units 0..63; 8 units = one 16-byte chunk
#pragma unroll
                for (int g = 0; g < UPT / 8; ++g)
                    *reinterpret_cast<uint4*>(sH + sw128_offset(row, sub * (UPT / 8) + g)) = make_uint4(pk[g][0], pk[g][1], pk[g][2], pk[g][3]);
            }
            fence_async_smem();                        // generic-proxy writes -> visible to tcgen05.mma / TMA
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_hdone);
            // ---------------- unit half 1
            mbar_wait(&d_full[1], par);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < UPT / 8; ++g)
                lstm_cell8<kFastAct, kTrain>(tbase + 256 + 8 * g, sbias + 256 + sub * UPT + 8 * g, c1s + 8 * g, pk[g],
                                             gd ? gd + 128 + 16 * g : nullptr, cd + 64 + 8 * g, hd + 64 + 8 * g);
            tc_fence_before();
            {
                uint8_t* sH = smem + SM_H + TILE;      // K-block 1 = units 64..127
#pragma unroll
                for (int g = 0; g < UPT / 8; ++g)
                    *reinterpret_cast<uint4*>(sH + sw128_offset(row, sub * (UPT / 8) + g)) = make_uint4(pk[g][0], pk[g][1], pk[g][2], pk[g][3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(leader_hdone + 8);
            named_bar(2, 32 * NEPI);
            if (storer) {
                tma_store_4d(&tmH, smem + SM_H, dir * 128, c1(t), c2(t), outer);
                tma_store_4d(&tmH, smem + SM_H + TILE, dir * 128 + 64, c1(t), c2(t), outer);
                bulk_commit();
            }
        }
        if (storer) bulk_wait0();
    }
    __syncwarp();

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc<2>(tmem, 512);
}

}  // namespace dprnn

using namespace dprnn;

// x [rows,128] bf16; w_packed [ndir*2*2*128, 256] bf16 (see engine._pack_lstm_tc); bias_perm [ndir,512] fp32;
// hout [rows, ndir*128] bf16.  Geometry: `inter`==0: rows = (b,s,k), sequences (b,s) run along k;
// `inter`==1: sequences (b,k) run along s.
static int lstm_layer_impl(const void* x, const void* w_packed, const float* bias_perm, void* hout, int B, int S, int K,
                           int inter, int hidden, int ndir, int fast_act, const int2* jobs, int n_jobs, void* stream,
                           void* gates = nullptr, float* cstate = nullptr, float* hout_f32 = nullptr) {
    DPRNN_CHECK_ARG(x && w_packed && bias_perm && hout && B > 0 && S > 0 && K > 0);
    DPRNN_CHECK_ARG(hidden == 128 && (ndir == 1 || ndir == 2));
    DPRNN_CHECK_ARG(((uintptr_t)x | (uintptr_t)w_packed | (uintptr_t)hout) % 16 == 0);
    CUtensorMap tmX, tmW, tmH;
    const uint64_t ldx = 128 * 2, ldh = (uint64_t)ndir * 128 * 2;
    LstmTcParams p;
    p.ndir = ndir;
    p.jobs = jobs;
    p.gates = (uint32_t*)gates; p.cst = cstate; p.hf = hout_f32;
    p.K = K; p.S = S;
    p.seq_limit = inter ? K : (long)B * S;
    uint64_t dX[4], sX[4], dH[4], sH[4];
    uint32_t box[4] = {64, 1, 1, 1};
    long njobs;
    if (!inter) {      // [feat, t=k (K), seq=(b,s) (B*S), 1]
        dX[0] = 128; dX[1] = K; dX[2] = (uint64_t)B * S; dX[3] = 1;
        sX[0] = 2; sX[1] = ldx; sX[2] = (uint64_t)K * ldx; sX[3] = (uint64_t)B * S * K * ldx;
        sH[0] = 2; sH[1] = ldh; sH[2] = (uint64_t)K * ldh; sH[3] = (uint64_t)B * S * K * ldh;
        box[2] = 128;
        p.T = K; p.seq_dim = 2; p.tiles_per_outer = (int)(((long)B * S + 255) / 256);
        njobs = (long)p.tiles_per_outer * ndir;
    } else {           // [feat, seq=k (K), t=s (S), b (B)]
        dX[0] = 128; dX[1] = K; dX[2] = S; dX[3] = B;
        sX[0] = 2; sX[1] = ldx; sX[2] = (uint64_t)K * ldx; sX[3] = (uint64_t)S * K * ldx;
        sH[0] = 2; sH[1] = ldh; sH[2] = (uint64_t)K * ldh; sH[3] = (uint64_t)S * K * ldh;
        box[1] = 128;
        p.T = S; p.seq_dim = 1; p.tiles_per_outer = (K + 255) / 256;
        njobs = (long)p.tiles_per_outer * B * ndir;
        if (jobs) {        // ragged: B == 1, S = total chunks of the packed batch, one pair-job per utterance and direction
            DPRNN_CHECK_ARG(B == 1 && K <= 256 && n_jobs > 0);
            njobs = (long)n_jobs * ndir;
        }
    }
    for (int i = 0; i < 4; ++i) dH[i] = dX[i];
    dH[0] = (uint64_t)ndir * 128;
    if (make_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dX, sX, box)) return 1;
    if (make_tmap(&tmH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, hout, dH, sH, box)) return 1;
    const uint64_t dW[2] = {256, (uint64_t)ndir * 512}, sW[2] = {2, 512};
    const uint32_t bW[2] = {64, 128};
    if (make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_packed, dW, sW, bW)) return 1;
    const size_t smem = SM_TOTAL;   // no alignment slack: the kernel checks the 1024-byte alignment of the base
    auto kern = gates ? (fast_act ? lstm_tc_kernel<true, true> : lstm_tc_kernel<false, true>)
                      : (fast_act ? lstm_tc_kernel<true, false> : lstm_tc_kernel<false, false>);
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DPRNN_CHECK_ARG(njobs * 2 < (1L << 31));
    kern<<<(unsigned)(njobs * 2), 128 + 32 * NEPI, smem, (cudaStream_t)stream>>>(tmX, tmW, tmH, bias_perm, p);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_lstm_layer_bf16(const void* x, const void* w_packed, const float* bias_perm, void* hout, int B,
                                     int S, int K, int inter, int hidden, int ndir, int fast_act, void* stream) {
    return lstm_layer_impl(x, w_packed, bias_perm, hout, B, S, K, inter, hidden, ndir, fast_act, nullptr, 0, stream);
}

// Training forward (cfg 5): the same kernel, whose epilogue also stores the gate activations (bf16, packed per 8-unit
// chunk: [rows][ndir][16][i8|f8|g8|o8]), the cell state and h (fp32) of every step - what dprnn_lstm_bptt_tc and the
// weight-gradient contractions read.
extern "C" int dprnn_lstm_layer_bf16_train(const void* x, const void* w_packed, const float* bias_perm, void* hout_bf16,
                                           void* gates, float* cstate, float* hout_f32, int B, int S, int K, int inter,
                                           int hidden, int ndir, int fast_act, void* stream) {
    DPRNN_CHECK_ARG(gates && cstate && hout_f32);
    DPRNN_CHECK_ARG(((uintptr_t)gates | (uintptr_t)cstate | (uintptr_t)hout_f32) % 16 == 0);
    return lstm_layer_impl(x, w_packed, bias_perm, hout_bf16, B, S, K, inter, hidden, ndir, fast_act, nullptr, 0, stream,
                           gates, cstate, hout_f32);
}

// Inter-chunk layer of a ragged batch: x / hout are the packed chunk space [total_chunks, K, .]; utt_jobs[j] =
// {first chunk, number of chunks} of one utterance (the caller orders them longest first: the pair-jobs are scheduled
// in that order, which is the LPT rule for the tail of the grid).
extern "C" int dprnn_lstm_inter_bf16_ragged(const void* x, const void* w_packed, const float* bias_perm, void* hout,
                                            long total_chunks, int K, const void* utt_jobs, int n_utt, int hidden,
                                            int ndir, int fast_act, void* stream) {
    DPRNN_CHECK_ARG(utt_jobs && n_utt > 0 && total_chunks > 0 && total_chunks < (1L << 31));
    return lstm_layer_impl(x, w_packed, bias_perm, hout, 1, (int)total_chunks, K, 1, hidden, ndir, fast_act,
                           (const int2*)utt_jobs, n_utt, stream);
}
