// Shared pieces of the fused tcgen05 LSTM kernels (lstm_tc.cu: one job per CTA pair; lstm_tc_sliced.cu: persistent,
// time-sliced jobs): shared-memory layout, parameters, small PTX helpers and the cell update.
#pragma once
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

constexpr int NEPI = 8;                       // epilogue warps: NEPI/4 per TMEM lane quadrant, 64/(NEPI/4) units per thread and half
constexpr int UPT = 64 / (NEPI / 4);          // units per thread and unit-half
constexpr int NXS = 4;                        // x ring stages, each a K-half [128 seq x 64 feat] bf16 = 16 KiB
constexpr uint32_t TILE = 128 * 128;          // bytes of one [128 rows x 128 B] swizzled tile
constexpr uint32_t SM_W = 0, SM_H = 8 * TILE, SM_X = SM_H + 2 * TILE, SM_BIAS = SM_X + NXS * TILE,
                   SM_BAR = SM_BIAS + 512 * 4, SM_TOTAL = SM_BAR + 128;
static_assert(SM_TOTAL <= 232448, "shared memory budget of one SM (227 KiB)");

struct LstmTcParams {
    int T;                 // time steps
    int seq_dim;           // which tensor-map coordinate runs over sequences: 2 (intra) or 1 (inter)
    int tiles_per_outer;   // 256-sequence tiles per outer index
    int ndir;
    const int2* jobs;      // ragged inter-chunk layer: per utterance {first chunk, number of chunks}; else NULL
    // training forward (cfg 5): what BPTT needs, stored by the epilogue as the values are produced
    uint32_t* gates;       // gate ACTIVATIONS i,f,g,o as bf16, packed [rows][ndir][16 chunks][4 gates][8 units]
    float* cst;            // [rows, ndir*H]  cell state after the step
    float* hf;             // [rows, ndir*H]  h in fp32 (operand of the weight-gradient contractions)
    int K, S;              // chunk length / chunks per utterance (linear row of (sequence, t))
    long seq_limit;        // number of real sequences along the sequence coordinate (tiles are padded to 256)
    int half_tiles;        // lstm_tc_pp_kernel: 128-sequence pair tiles = half-job A only (small batches: twice the pairs busy)
    // fused input norm (lstm_tc_pp_kernel<kFuse>): the layer's input is x_in + norm(y) of the PREVIOUS half-block's tail,
    // applied to every x tile in shared memory before the tensor core reads it; direction 0 also writes it to x_out
    const uint4* fy;       // [rows, 128] 16-bit: Linear output of the previous half-block
    const float2* fmr;     // [B] {mean, rstd} of that output per utterance
    const float *fgamma, *fbeta;     // [128] affine of the norm
    uint4* fxout;          // [rows, 128] 16-bit: the updated residual stream (a buffer other than the x the layer reads)
};

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
// x tile load whose completion is signalled on the LEADER CTA's mbarrier (cta_group::2 form)
__device__ __forceinline__ void tma_load_4d_pair(void* smem, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1,
                                                 int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(m), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(m), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// Remote arrive with the default (.release.cta) semantics, as CUTLASS' ClusterBarrier::arrive does: the data the
// barrier guards (the h tile) is read by the tensor core through the async proxy and has already been published with
// fence.proxy.async; a .release.cluster here compiles to MEMBAR.ALL.GPU + CCTL.IVALL and cost ~19% of the epilogue.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}

// 256-bit store (STG.256): 8 consecutive floats = one full 32-byte sector per thread
__device__ __forceinline__ void st_global_v8(float* p, const float (&a)[4], const float (&b)[4]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(b[0]), "f"(b[1]), "f"(b[2]), "f"(b[3]) : "memory");
}

__device__ __forceinline__ void st_global_v8u(uint32_t* p, const uint32_t (&a)[4], const uint32_t (&b)[4]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// Cell update for 8 hidden units of one sequence row: gates from TMEM (+bias) -> c (registers), h (packed bf16, one
// 16-byte chunk).  8 units at a time keeps the live register set small enough for 128 registers per thread, which
// leaves room on the SM for a memory-bound CTA of another stream next to this kernel.
// The i, f, o rows of the weights and biases are pre-scaled by 1/2 on the host (exact), so sigmoid(x) = 1/2 tanh(x') + 1/2.
// kGS = distance in TMEM columns between the four gates of a unit (64: one-job kernels; 32: the half-job kernel).
// kF16: h leaves as fp16 instead of bf16 (the 'fp16' mode: operands with 11 significand bits).
// kStage (training, 128-sequence tiles): the saved values do not go to global memory from the registers (32 rows x 32 B
// per store instruction = 32 L1 wavefronts of a quarter line each; the step was bound by exactly that) but into a
// warp-private staging buffer `stg` laid out as three TMA boxes - gates [32 rows x 64 B] SWIZZLE_64B at 0, c and h
// [32 rows x 32 B] SWIZZLE_32B at 2048 / 3072 - which the caller stores with three bulk tensor copies.
template <bool kFastAct, bool kTrain, int kGS = 64, bool kF16 = false, bool kStage = false>
__device__ __forceinline__ void lstm_cell8(uint32_t tcol, const float* __restrict__ bq, float* __restrict__ c,
                                           uint32_t (&packed)[4], uint32_t* __restrict__ gdst = nullptr,
                                           float* __restrict__ cdst = nullptr, float* __restrict__ hdst = nullptr,
                                           uint8_t* __restrict__ stg = nullptr, int lane = 0) {
    uint32_t ri[8], rf[8], rg[8], ro[8];
    tmem_ld8_issue(tcol + 0 * kGS, ri);
    tmem_ld8_issue(tcol + 1 * kGS, rf);
    tmem_ld8_issue(tcol + 2 * kGS, rg);
    tmem_ld8_issue(tcol + 3 * kGS, ro);
    tmem_ld_wait();
    float keep[kTrain ? 2 : 1][4];
    uint32_t gk[kTrain ? 4 : 1][4];      // bf16 pairs of the 4 gates of the call's 8 units
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
        const float4 bi = *reinterpret_cast<const float4*>(bq + 0 * 64 + j), bf = *reinterpret_cast<const float4*>(bq + 1 * 64 + j);
        const float4 bg = *reinterpret_cast<const float4*>(bq + 2 * 64 + j), bo = *reinterpret_cast<const float4*>(bq + 3 * 64 + j);
        const float b4[4][4] = {{bi.x, bi.y, bi.z, bi.w}, {bf.x, bf.y, bf.z, bf.w}, {bg.x, bg.y, bg.z, bg.w}, {bo.x, bo.y, bo.z, bo.w}};
        float hv[4], sv[5][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float pi = __uint_as_float(ri[j + u]) + b4[0][u];
            const float pf = __uint_as_float(rf[j + u]) + b4[1][u];
            const float pg = __uint_as_float(rg[j + u]) + b4[2][u];
            const float po = __uint_as_float(ro[j + u]) + b4[3][u];
            float ig, fg, gg, og;
            if constexpr (kFastAct) {
                ig = fmaf(tanh_fast(pi), 0.5f, 0.5f); fg = fmaf(tanh_fast(pf), 0.5f, 0.5f);
                gg = tanh_fast(pg); og = fmaf(tanh_fast(po), 0.5f, 0.5f);
            } else {
                ig = sigmoid_acc(2.f * pi); fg = sigmoid_acc(2.f * pf); gg = tanhf(pg); og = sigmoid_acc(2.f * po);
            }
            const float cn = fmaf(fg, c[j + u], ig * gg);
            c[j + u] = cn;
            hv[u] = og * (kFastAct ? tanh_fast(cn) : tanhf(cn));
            if constexpr (kTrain) { sv[0][u] = ig; sv[1][u] = fg; sv[2][u] = gg; sv[3][u] = og; sv[4][u] = cn; }
        }
        if constexpr (kTrain && kStage) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                gk[g][(j >> 1) + 0] = pack_bf16x2(sv[g][0], sv[g][1]);
                gk[g][(j >> 1) + 1] = pack_bf16x2(sv[g][2], sv[g][3]);
            }
            // c / h: 16-byte chunk j/4 of the row's 32 bytes, SWIZZLE_32B (chunk ^= bit 7 of the byte offset = bit 2 of the row)
            const uint32_t o32 = (uint32_t)lane * 32 + (uint32_t)(((j >> 2) ^ ((lane >> 2) & 1)) << 4);
            *reinterpret_cast<float4*>(stg + 2048 + o32) = make_float4(sv[4][0], sv[4][1], sv[4][2], sv[4][3]);
            if (hdst) *reinterpret_cast<float4*>(stg + 3072 + o32) = make_float4(hv[0], hv[1], hv[2], hv[3]);   // hdst: only "h wanted"
            if (j == 4) {
                // gates: chunk g (8 units of gate g) of the row's 64 bytes, SWIZZLE_64B (chunk ^= bits 7..8 = bits 1..2 of the row)
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<uint4*>(stg + (uint32_t)lane * 64 + (uint32_t)((g ^ ((lane >> 1) & 3)) << 4)) =
                        make_uint4(gk[g][0], gk[g][1], gk[g][2], gk[g][3]);
            }
        } else if constexpr (kTrain) {
            if (gdst) {
                // the 8 units of the call leave as 64 contiguous bytes of bf16 gates [i8 | f8 | g8 | o8] (two 32-byte
                // stores) plus one 32-byte store each for c and h (fp32): 4 full-sector stores per thread and call
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    gk[g][(j >> 1) + 0] = pack_bf16x2(sv[g][0], sv[g][1]);
                    gk[g][(j >> 1) + 1] = pack_bf16x2(sv[g][2], sv[g][3]);
                }
                if (j == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) { keep[0][u] = sv[4][u]; keep[1][u] = hv[u]; }
                } else {
                    st_global_v8u(gdst, gk[0], gk[1]);
                    st_global_v8u(gdst + 8, gk[2], gk[3]);
                    st_global_v8(cdst, keep[0], sv[4]);
                    if (hdst) st_global_v8(hdst, keep[1], hv);
                }
            }
        }
        packed[(j >> 1) + 0] = pack_h16x2<kF16>(hv[0], hv[1]);
        packed[(j >> 1) + 1] = pack_h16x2<kF16>(hv[2], hv[3]);
    }
}


}  // namespace dprnn
