// BPTT of one LSTM layer with the recurrent contraction on 5th-gen tensor cores (training step, tensor-core mode):
//     d h_{t-1}[256 seq, 128] = d gates_t[256 seq, 512] (bf16) @ W_hh[512, 128] (bf16)      (fp32 accumulators in TMEM)
// Inputs: d h_out (fp32, row-major), the gate activations saved by the tensor-core forward (bf16, packed per 8-unit chunk)
// and the cell states (fp32); output: the gradient of the gate pre-activations (fp32, row-major), as the exact fp32 kernel
// (lstm_bwd_kernel, lstm_simt.cu) produces it.
//
// A CTA PAIR (cluster of 2, tcgen05 cta_group::2, M = 256) owns 256 sequences of one direction for all T steps; each CTA
// holds its 128 sequences' d gates tile (A operand, 8 K-blocks of [128 x 64] bf16, 128 KiB, rewritten every step) and
// half of W_hh^T (B operand, 64 of the 128 output columns x K = 512, 64 KiB, resident).
//
// The element-wise part keeps the COALESCED thread mapping of the SIMT kernel (half-warp = one sequence row, lane = 4
// consecutive hidden units), because the step streams 5.5 KB per row through it (4 gates, c_t, c_{t-1}, d h_out in,
// 4 d gates out) and a thread-per-row mapping would touch 32 cache lines per load instruction.  TMEM, however, is read
// thread-per-row (tcgen05.ld 32x32b), so the recurrent d h goes TMEM -> registers -> padded, warp-private staging rows
// ([128 rows x 64 units] fp32, one unit-half at a time) -> the coalesced mapping.  Per step:
//     wait D[prev] | for unit-half p in {0,1}: stage D[prev][:, 64p..64p+63]; cell backward for those units (all 128 rows):
//     d gates -> global (fp32) and -> the A tile (bf16, 128B-swizzled K-major); publish -> the MMA warp issues the K-blocks of
//     that unit-half (the MMAs of half 0 overlap the element-wise work of half 1) | commit -> D[cur]   (TMEM ping-pong)
// K index = gate*128 + unit, i.e. K-block kb = 2*gate + unit_half: W_hh^T needs no permutation.
//
// kOutBf16: d gates leaves as bf16 [rows, ndir*4H] - exactly the A tile the step has just built for the tensor core: one
// warp stores its 8 K-blocks with TMA (box = [rows of the tile x 64 columns], rows past the last sequence are clipped)
// instead of sixteen warps issuing fp32 vector stores.  Half the bytes for this kernel and for the three passes that read
// d gates afterwards (d x and the weight gradients, which consume bf16 operands: gemm_kdeep.cu, atb_tc.cu).
//
// kRows = 64 (small batches: at 16 utterances per GPU the 256-sequence tiles give 52..64 CTAs for 148 SMs, and the step is
// bound by the latency of the streamed loads, not by their bandwidth): the pair owns 128 sequences, 64 per CTA, and the
// MMA is the cta_group::2 M = 128 shape.  Its accumulator has the "2x2" layout - TMEM lanes 0..63 hold units 0..63 of the
// CTA's 64 rows, lanes 64..127 hold units 64..127 of the SAME rows (64 columns per buffer) - so the lane quadrant a warp
// may read fixes both its rows and its unit-half: every warp differentiates one unit-half of 8 rows per step instead of
// both halves, twice as many SMs work, and c_{t-1} of a step is kept in registers as c_t of the next one.
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

namespace bptt {
constexpr int H = 128, G4 = 512;
constexpr uint32_t A_TILE = 128 * 128;        // [128 rows x 128 B] (kRows = 64: the first half of each tile is used)
constexpr uint32_t W_TILE = 64 * 128;         // [64 rows x 128 B]
constexpr int NEW = 16;                       // element-wise warps (issue-bound part: more warps hide its latencies)
constexpr int ITS = 128 / NEW / 2;            // row pairs per warp and unit-half
constexpr int STG_LD = 68;                    // floats per staging row (64 + 4: conflict-free float4 rows)
constexpr uint32_t SM_W = 0, SM_A = 8 * W_TILE, SM_STG = SM_A + 8 * A_TILE, SM_BAR = SM_STG + 128 * STG_LD * 4,
                   SM_TOTAL = SM_BAR + 128;
static_assert(SM_TOTAL <= 232448, "shared memory budget of one SM (227 KiB)");
}  // namespace bptt

struct LstmBpttTcParams {
    const float* dh_out;   // [rows, ndir*H]
    const uint2* gates;    // activations i,f,g,o as bf16, packed [rows][ndir][16 chunks of 8 units][4 gates][8] (the
                           // layout dprnn_lstm_layer_bf16_train writes); one uint2 = 4 units of one gate
    const float* cstate;   // [rows, ndir*H]
    float* dgates;         // [rows, ndir*4H] fp32 (kOutBf16: unused - the bf16 output leaves through the tensor map)
    int T;
    // sequences are tiled per OUTER index (intra-chunk layer: one outer, the B*S sequences; inter-chunk: one outer per
    // utterance, its K sequences), so that a tile's rows are one box of the output tensor map; row of (outer, n, t) =
    // outer * seq_outer + n * seq_inner + t * step_stride
    int limit, tiles_per_outer;
    long seq_outer, seq_inner, step_stride;
    int ndir;
};

__device__ __forceinline__ float bptt_tanh(float x, bool fast) {
    if (fast) {
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    return tanhf(x);
}
// d gates is written once and read back much later (3.2 GB per layer): streaming stores keep it from evicting the
// L2-prefetched inputs of the next step
__device__ __forceinline__ void st_stream4(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t bptt_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void bptt_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t bptt_map_to_cta(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void bptt_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void bptt_tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(m), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bptt_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <bool kFastAct, int kRows, bool kOutBf16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * bptt::NEW, 1)
lstm_bptt_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmDG,
                    const LstmBpttTcParams p) {
    using namespace bptt;
    static_assert(kRows == 128 || kRows == 64, "rows per CTA");
    constexpr int NPH = kRows == 128 ? 2 : 1;          // unit-halves a warp differentiates
    constexpr uint32_t DCOLS = kRows == 128 ? 128 : 64; // TMEM columns of one accumulator buffer
    constexpr bool kKeepC = kRows == 64;               // c_{t-1} stays in registers for the next step
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* w_full = bars;
    uint64_t* d_full = bars + 1;              // MMAs of the step complete (multicast commit, both CTAs)
    uint64_t* a_ready = bars + 2;             // [2] (leader's copy) A K-blocks of unit-half p written by both CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    uint64_t* a_local = bars + 5;             // [2] kOutBf16: this CTA's warps have written unit-half p of the A tile
    uint64_t* a_free = bars + 7;              // kOutBf16: the bulk stores of the step have read the A tile
    float* stg = reinterpret_cast<float*>(smem + SM_STG);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = bptt_cluster_ctarank();
    const int job = blockIdx.x >> 1;
    const int dir = job % p.ndir;
    const int jt = job / p.ndir;
    const int outer = jt / p.tiles_per_outer;
    const int n0 = (jt % p.tiles_per_outer) * (2 * kRows) + (int)rank * kRows;      // first sequence of this CTA inside `outer`
    const int T = p.T;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) {
            printf("lstm_bptt_tc_kernel: dynamic shared memory base %u is not 1024-byte aligned\n", smem_u32(smem));
            __trap();
        }
        prefetch_tmap(&tmW);
        mbar_init(w_full, 1);
        mbar_init(d_full, 1);
        // one elected lane per element-wise warp that writes the unit-half, both CTAs
        mbar_init(&a_ready[0], 2 * NEW / (3 - NPH)); mbar_init(&a_ready[1], 2 * NEW / (3 - NPH));
        mbar_init(&a_local[0], NEW / (3 - NPH)); mbar_init(&a_local[1], NEW / (3 - NPH));
        mbar_init(a_free, 1);
        if constexpr (kOutBf16) prefetch_tmap(&tmDG);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<2>(tmem_slot, 256);
    tc_fence_before();
    bptt_cluster_sync();
    tc_fence_after();

    if (warp == 0 && elect_one()) {           // this CTA's 64 output columns of W_hh^T, all 8 K-blocks
        mbar_expect_tx(w_full, 8 * W_TILE);
        for (int kb = 0; kb < 8; ++kb)
            tma_load_2d(smem + SM_W + kb * W_TILE, &tmW, w_full, kb * 64, dir * 128 + (int)rank * 64);
    }
    mbar_wait(w_full, 0);
    bptt_cluster_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * kRows, 128);
            const uint32_t aW = smem_u32(smem + SM_W), aA = smem_u32(smem + SM_A);
            for (int s = 0; s + 1 < T; ++s) {
                const uint32_t d = tmem + (uint32_t)(s & 1) * DCOLS;
                for (int half = 0; half < 2; ++half) {
                    mbar_wait_cluster(&a_ready[half], s & 1);
                    tc_fence_after();
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int kb = q4 * 2 + half;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16<2>(d, umma_desc_sw128(aA + kb * A_TILE + kk * 32),
                                         umma_desc_sw128(aW + kb * W_TILE + kk * 32), idesc,
                                         (half | q4 | kk) ? 1u : 0u);
                    }
                }
                umma_commit_2cta(d_full, 3);
            }
        }
        __syncwarp();
    } else if (kOutBf16 && warp == 3) {
        // ================= d gates of the step: the A tile, as it is, to global memory (bf16) =================
        if (elect_one()) {
            for (int s = 0; s < T; ++s) {
                const int fstep = T - 1 - s;
                const int t = dir ? T - 1 - fstep : fstep;
                for (int half = 0; half < 2; ++half) {
                    mbar_wait(&a_local[half], s & 1);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int kb = q4 * 2 + half;
                        bptt_tma_store_4d(&tmDG, smem + SM_A + kb * A_TILE, dir * G4 + kb * 64, n0, t, outer);
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(a_free);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ================= element-wise backward of the cell =================
        // warp -> TMEM lane quadrant q (= warp % 4, the only lanes it may read) and the wq-th group of 2*ITS rows inside it:
        // the rows a warp differentiates are rows whose recurrent d h it can fetch itself, so staging is warp-private
        const int e = warp - 4, q = e & 3, wq = e >> 2;
        const int hw = lane >> 4, l16 = lane & 15;           // coalesced role: half-warp = row, 4 units per lane
        const int rq = (kRows == 128 ? q : (q & 1)) * 32 + wq * (2 * ITS);    // first of this warp's rows inside the CTA's tile
        const int ph0 = kRows == 128 ? 0 : (q >> 1);         // first (kRows = 64: only) unit-half of this warp
        const int ldg = p.ndir * G4, ldh = p.ndir * H;
        const uint32_t leader_ready = bptt_map_to_cta(smem_u32(&a_ready[0]), 0);
        float* stgw = stg + e * (2 * ITS) * STG_LD;          // this warp's staging rows
        int base[ITS];
        bool ok[ITS];
#pragma unroll
        for (int it = 0; it < ITS; ++it) {
            const int n = n0 + rq + it * 2 + hw;
            ok[it] = n < p.limit;
            base[it] = ok[it] ? (int)((long)outer * p.seq_outer + (long)n * p.seq_inner) : 0;
        }
        float dc[NPH][ITS][4];
        float4 ckeep[kKeepC ? ITS : 1];                      // kKeepC: c_{t-1} of this step = c_t of the next
#pragma unroll
        for (int a = 0; a < NPH; ++a)
#pragma unroll
            for (int b = 0; b < ITS; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) dc[a][b][c] = 0.f;

        // The step streams 7 float4 per (row, lane) through registers; with only 8 warps per SM the loads of the NEXT
        // (row pair | unit-half | step) are issued before the current one is computed (software double buffering).
        struct Ld { uint2 gi, gf, gg, go; float4 cv, cp, dho; };
        // Row indices fit 32 bits (checked by the launcher): every address is base + (u32 row) * (u32 stride) + constant,
        // one IMAD.WIDE each instead of 64-bit multiply chains (address arithmetic was 37 % of the executed instructions).
        const int sstep = (int)p.step_stride;
        const unsigned gstride = (unsigned)p.ndir * 128u, hstride = (unsigned)ldh;      // uint2 units / floats per row
        auto gate_off = [&](int ph_) { const int u0 = ph_ * 64 + l16 * 4; return (unsigned)(dir * 128 + (u0 >> 3) * 8 + ((u0 >> 2) & 1)); };
        auto ch_off = [&](int ph_) { return (unsigned)(dir * H + ph_ * 64 + l16 * 4); };
        auto issue = [&](int s_, int ph_, bool valid, int base_, Ld& L) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const uint2 zu = make_uint2(0u, 0u);
            L.gi = zu; L.gf = zu; L.gg = zu; L.go = zu; L.cv = z; L.cp = z; L.dho = z;
            if (s_ < T && valid) {
                const int t_ = dir ? s_ : T - 1 - s_;          // time of the forward step T-1-s_ in this direction's order
                const unsigned R = (unsigned)(base_ + t_ * sstep), Rp = dir ? R + (unsigned)sstep : R - (unsigned)sstep;
                // chunk = u0 / 8 (8 uint2 each: 2 per gate), 4-unit half (u0 / 4) & 1 inside it
                const uint2* g = p.gates + (size_t)R * gstride + gate_off(ph_);
                L.gi = __ldg(g); L.gf = __ldg(g + 2); L.gg = __ldg(g + 4); L.go = __ldg(g + 6);
                const unsigned co = ch_off(ph_);
                if (!kKeepC || s_ == 0) L.cv = *reinterpret_cast<const float4*>(p.cstate + (size_t)R * hstride + co);
                if (s_ < T - 1) L.cp = *reinterpret_cast<const float4*>(p.cstate + (size_t)Rp * hstride + co);
                L.dho = ld_stream(reinterpret_cast<const float4*>(p.dh_out + (size_t)R * hstride + co));
            }
        };
        // L2 prefetch of the step after the next one's register prefetch can reach: the addresses of every step are known up
        // front, so DRAM latency is taken off the recurrence's critical path (two lanes per half-warp cover its two lines)
        auto prefetch_step = [&](int s_) {
            if (s_ >= T || (l16 & 7) != 0) return;
            const int t_ = dir ? s_ : T - 1 - s_;
#pragma unroll
            for (int it = 0; it < ITS; ++it) {
                if (!ok[it]) continue;
                const unsigned R = (unsigned)(base[it] + t_ * sstep), Rp = dir ? R + (unsigned)sstep : R - (unsigned)sstep;
#pragma unroll
                for (int pi_ = 0; pi_ < NPH; ++pi_) {
                    // packed gates: 512 B = 4 lines per (row, direction, unit-half); lanes 0 and 8 of the half-warp fetch two
                    // consecutive lines each (chunks 0-3 / 4-7)
                    const uint2* g = p.gates + (size_t)R * gstride + (gate_off(ph0 + pi_) & ~7u);
                    const unsigned co = ch_off(ph0 + pi_);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(g + 16));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dh_out + (size_t)R * hstride + co));
                    if (s_ < T - 1)     // c_{t-1} of that step (it is c_t of the step after it)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.cstate + (size_t)Rp * hstride + co));
                }
            }
        };
        Ld cur, nxt;
        issue(0, ph0, ok[0], base[0], cur);

        for (int s = 0; s < T; ++s) {
            const int fstep = T - 1 - s;                      // forward step being differentiated
            const int t = dir ? T - 1 - fstep : fstep;
            prefetch_step(s + 1);
            if (s > 0) {
                mbar_wait(d_full, (s - 1) & 1);
                tc_fence_after();
                if constexpr (kOutBf16) mbar_wait(a_free, (s - 1) & 1);      // ... and its bulk stores have read the A tile
            }
#pragma unroll
            for (int pi = 0; pi < NPH; ++pi) {
                const int ph = ph0 + pi;
                if (s > 0) {                                  // recurrent d h of units 64ph..64ph+63: TMEM -> staging
                    __syncwarp();                             // the previous unit-half's reads of the staging rows are done
                    const bool mine = lane >= wq * (2 * ITS) && lane < (wq + 1) * (2 * ITS);
                    float* dst = stgw + (lane - wq * (2 * ITS)) * STG_LD;
#pragma unroll
                    for (int cq = 0; cq < 4; ++cq) {
                        float v[16];
                        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((s - 1) & 1) * DCOLS + pi * 64 + cq * 16, v);
                        if (mine) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                *reinterpret_cast<float4*>(dst + cq * 16 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                }
                const int u0 = ph * 64 + l16 * 4;
#pragma unroll
                for (int it = 0; it < ITS; ++it) {
                    // next in the order (step, unit-half, row pair)
                    if (it < ITS - 1) issue(s, ph, ok[it + 1], base[it + 1], nxt);
                    else if (pi < NPH - 1) issue(s, 1, ok[0], base[0], nxt);
                    else issue(s + 1, ph0, ok[0], base[0], nxt);
                    if constexpr (kKeepC) {
                        if (s > 0) cur.cv = ckeep[it];
                        ckeep[it] = cur.cp;
                    }
                    const int row = rq + it * 2 + hw;
                    const unsigned rowi = (unsigned)(base[it] + t * sstep);
                    float4 dhr = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (s > 0) dhr = *reinterpret_cast<const float4*>(stgw + (it * 2 + hw) * STG_LD + l16 * 4);
                    auto unpack = [](uint2 v, float (&o)[4]) {
                        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
                        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
                        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
                    };
                    float ia[4], fa[4], ga[4], oa[4];
                    unpack(cur.gi, ia); unpack(cur.gf, fa); unpack(cur.gg, ga); unpack(cur.go, oa);
                    const float ca[4] = {cur.cv.x, cur.cv.y, cur.cv.z, cur.cv.w}, pa[4] = {cur.cp.x, cur.cp.y, cur.cp.z, cur.cp.w},
                                da[4] = {cur.dho.x + dhr.x, cur.dho.y + dhr.y, cur.dho.z + dhr.z, cur.dho.w + dhr.w};
                    float dpi[4], dpf[4], dpg[4], dpo[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float dh = da[u];
                        const float tcv = bptt_tanh(ca[u], kFastAct);
                        const float dct = dc[pi][it][u] + dh * oa[u] * (1.f - tcv * tcv);
                        dpo[u] = dh * tcv * oa[u] * (1.f - oa[u]);
                        dpi[u] = dct * ga[u] * ia[u] * (1.f - ia[u]);
                        dpf[u] = dct * pa[u] * fa[u] * (1.f - fa[u]);
                        dpg[u] = dct * ia[u] * (1.f - ga[u] * ga[u]);
                        dc[pi][it][u] = dct * fa[u];
                    }
                    if (!kOutBf16 && ok[it]) {
                        float* o = p.dgates + (size_t)rowi * (unsigned)ldg + (unsigned)(dir * G4 + u0);
                        st_stream4(o, dpi[0], dpi[1], dpi[2], dpi[3]);
                        st_stream4(o + H, dpf[0], dpf[1], dpf[2], dpf[3]);
                        st_stream4(o + 2 * H, dpg[0], dpg[1], dpg[2], dpg[3]);
                        st_stream4(o + 3 * H, dpo[0], dpo[1], dpo[2], dpo[3]);
                    }
                    if (kOutBf16 || fstep > 0) {
                        // A operand: K-block kb = 2*gate + ph, column (unit % 64) -> 16-byte chunk l16/2, byte (l16&1)*8
                        const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)(l16 >> 1)) + (uint32_t)(l16 & 1) * 8;
                        uint8_t* at = smem + SM_A + ph * A_TILE + off;
                        auto pack = [](float a, float b) {
                            __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
                            return *reinterpret_cast<uint32_t*>(&v);
                        };
                        *reinterpret_cast<uint2*>(at + 0 * 2 * A_TILE) = make_uint2(pack(dpi[0], dpi[1]), pack(dpi[2], dpi[3]));
                        *reinterpret_cast<uint2*>(at + 1 * 2 * A_TILE) = make_uint2(pack(dpf[0], dpf[1]), pack(dpf[2], dpf[3]));
                        *reinterpret_cast<uint2*>(at + 2 * 2 * A_TILE) = make_uint2(pack(dpg[0], dpg[1]), pack(dpg[2], dpg[3]));
                        *reinterpret_cast<uint2*>(at + 3 * 2 * A_TILE) = make_uint2(pack(dpo[0], dpo[1]), pack(dpo[2], dpo[3]));
                    }
                    cur = nxt;
                }
                if (kOutBf16 || fstep > 0) {
                    fence_async_smem();                       // generic-proxy writes of the A tile -> visible to tcgen05.mma / TMA
                    __syncwarp();
                    if (lane == 0) {
                        if (fstep > 0) bptt_arrive_remote(leader_ready + ph * 8);
                        if constexpr (kOutBf16) mbar_arrive(&a_local[ph]);
                    }
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    bptt_cluster_sync();
    if (warp == 2) tmem_dealloc<2>(tmem, 256);
}

}  // namespace dprnn

using namespace dprnn;

// whhT_bf16: [ndir][H = 128 output columns j][4H = 512 gate rows k] bf16 = W_hh^T per direction (k = gate*128 + unit)
static int bptt_tc_impl(const float* dh_out, const void* gates, const float* cstate, const void* whhT_bf16, float* dgates,
                        void* dgates_bf16, long nseq, int T, long seq_div, long seq_outer_stride, long seq_inner_stride,
                        long step_stride, int hidden, int ndir, int fast_act, void* stream) {
    DPRNN_CHECK_ARG(dh_out && gates && cstate && whhT_bf16 && (dgates || dgates_bf16) && nseq > 0 && T > 0 && seq_div > 0);
    DPRNN_CHECK_ARG(hidden == 128 && (ndir == 1 || ndir == 2) && seq_div < (1L << 31) && nseq < (1L << 31));
    DPRNN_CHECK_ARG(((uintptr_t)dh_out | (uintptr_t)gates | (uintptr_t)cstate | (uintptr_t)whhT_bf16 | (uintptr_t)dgates |
                     (uintptr_t)dgates_bf16) % 16 == 0);
    // row indices are kept as int32 inside the kernel
    const long max_row = ((nseq - 1) / seq_div) * seq_outer_stride + ((nseq - 1) % seq_div) * seq_inner_stride +
                         (long)(T - 1) * step_stride;
    DPRNN_CHECK_ARG(max_row < (1L << 31));
    // tiles per outer index: seq_div == 1 -> one outer holding all sequences (row = n * seq_outer_stride + t * step_stride),
    // else nseq / seq_div outers of seq_div sequences each
    DPRNN_CHECK_ARG(nseq % seq_div == 0);
    const long limit = seq_div == 1 ? nseq : seq_div, n_outer = seq_div == 1 ? 1 : nseq / seq_div;
    const long sn = seq_div == 1 ? seq_outer_stride : seq_inner_stride, so = seq_div == 1 ? 0 : seq_outer_stride;
    CUtensorMap tmW, tmDG;
    const uint64_t dW[2] = {512, (uint64_t)ndir * 128}, sW[2] = {2, 1024};
    const uint32_t bW[2] = {64, 64};
    if (make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, whhT_bf16, dW, sW, bW)) return 1;
    // small batches: 64 rows per CTA when the 128-sequence pair tiles still fit the CTA pairs in one wave (the rule of the
    // forward kernel, lstm_tc_pp.cu); DPRNN_LSTM_HALF_TILES / DPRNN_LSTM_FULL_TILES in `fast_act` force the choice
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long jobs128 = (limit + 127) / 128 * n_outer * ndir;
    const bool half = (fast_act & DPRNN_LSTM_HALF_TILES) || (!(fast_act & DPRNN_LSTM_FULL_TILES) && jobs128 <= sms / 2);
    const int tile = half ? 128 : 256;
    const long tpo = (limit + tile - 1) / tile, njobs = tpo * n_outer * ndir;
    DPRNN_CHECK_ARG(njobs * 2 < (1L << 31));
    tmDG = tmW;
    if (dgates_bf16) {      // [rows, ndir*512] bf16 addressed as (column, sequence inside the outer, time, outer)
        const uint64_t ld = (uint64_t)ndir * 512 * 2;
        const uint64_t dG[4] = {(uint64_t)ndir * 512, (uint64_t)limit, (uint64_t)T, (uint64_t)n_outer};
        const uint64_t sG[4] = {2, (uint64_t)sn * ld, (uint64_t)step_stride * ld, (uint64_t)(so ? so : 1) * ld};
        const uint32_t bG[4] = {64, (uint32_t)(tile / 2), 1, 1};
        if (make_tmap(&tmDG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dgates_bf16, dG, sG, bG)) return 1;
    }
    LstmBpttTcParams p{dh_out, (const uint2*)gates, cstate, dgates, T, (int)limit, (int)tpo, so, sn, step_stride, ndir};
    const bool fa = fast_act & DPRNN_LSTM_FAST_ACT, ob = dgates_bf16 != nullptr;
    auto kern = half ? (ob ? (fa ? lstm_bptt_tc_kernel<true, 64, true> : lstm_bptt_tc_kernel<false, 64, true>)
                           : (fa ? lstm_bptt_tc_kernel<true, 64, false> : lstm_bptt_tc_kernel<false, 64, false>))
                     : (ob ? (fa ? lstm_bptt_tc_kernel<true, 128, true> : lstm_bptt_tc_kernel<false, 128, true>)
                           : (fa ? lstm_bptt_tc_kernel<true, 128, false> : lstm_bptt_tc_kernel<false, 128, false>));
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bptt::SM_TOTAL));
    kern<<<(unsigned)(njobs * 2), 128 + 32 * bptt::NEW, bptt::SM_TOTAL, (cudaStream_t)stream>>>(tmW, tmDG, p);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_lstm_bptt_tc(const float* dh_out, const void* gates, const float* cstate, const void* whhT_bf16,
                                  float* dgates, long nseq, int T, long seq_div, long seq_outer_stride,
                                  long seq_inner_stride, long step_stride, int hidden, int ndir, int fast_act,
                                  void* stream) {
    DPRNN_CHECK_ARG(dgates);
    return bptt_tc_impl(dh_out, gates, cstate, whhT_bf16, dgates, nullptr, nseq, T, seq_div, seq_outer_stride,
                        seq_inner_stride, step_stride, hidden, ndir, fast_act, stream);
}

// d gates as bf16 [rows, ndir*4H] (stored by TMA from the tile the tensor core reads): the operand format of
// dprnn_gemm_kdeep (d x) and dprnn_gemm_atb_dual (weight gradients) in their bf16 forms
extern "C" int dprnn_lstm_bptt_tc_bf16out(const float* dh_out, const void* gates, const float* cstate, const void* whhT_bf16,
                                          void* dgates_bf16, long nseq, int T, long seq_div, long seq_outer_stride,
                                          long seq_inner_stride, long step_stride, int hidden, int ndir, int fast_act,
                                          void* stream) {
    DPRNN_CHECK_ARG(dgates_bf16);
    return bptt_tc_impl(dh_out, gates, cstate, whhT_bf16, nullptr, dgates_bf16, nseq, T, seq_div, seq_outer_stride,
                        seq_inner_stride, step_stride, hidden, ndir, fast_act, stream);
}
