// Persistent, fully pipelined bf16 tensor-core Linear for the layer that follows every LSTM:
//     C[M,128] (fp32) = A[M,K] (bf16) @ W[128,K]^T (bf16) + bias,   K in {128, 256}
// plus per-row {sum, sumsq} of the output for the GroupNorm / gLN that follows.  The op is HBM-bound
// (read 256-512 B, write 512 B per row, 64-128 KFLOP), so the kernel is built to keep TMA busy:
//   * one CTA per SM loops over 128-row tiles; W (32-64 KiB) is loaded once and stays in shared memory;
//   * A K-blocks (16 KiB) stream through a 6-stage TMA ring;
//   * two 128-column TMEM accumulators: the MMA of tile i+1 overlaps the epilogue of tile i;
//   * epilogue: tcgen05.ld -> +bias -> 128B-swizzled shared-memory staging (conflict-free) -> TMA store, so
//     HBM sees full 128-byte lines instead of per-thread 16-byte fragments.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

constexpr int LP_N = 128, LP_AST = 6, LP_CST = 2, LP_TQ = 4;
constexpr uint32_t LP_BLK = 128 * 128;                 // one [128 rows x 128 B] swizzled tile = 16 KiB

struct LinPersistArgs {
    const float* bias;
    float2* stats;        // [M] per-row {sum, sumsq}, or NULL
    int M, tiles;
    unsigned* ticket;     // zeroed before launch: tiles are handed out dynamically (NULL: static round-robin).  With
                          // one CTA per SM and other kernels sharing the GPU (utterance groups on other streams), some
                          // CTAs start late; a static partition would make the whole launch wait for them.
};

__device__ __forceinline__ void lp_tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(m), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
}

template <int KB, bool kBf16Out, bool kF16>     // K / 64; output fp32 or 16-bit; operands (and 16-bit output) bf16 / fp16
__global__ void __launch_bounds__(192, 1) linear_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmW,
                                                                const __grid_constant__ CUtensorMap tmC,
                                                                const LinPersistArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;                                 // KB blocks
    uint8_t* sA = sW + KB * LP_BLK;                     // LP_AST blocks
    uint8_t* sC = sA + LP_AST * LP_BLK;                 // LP_CST staging blocks [128 rows x 32 fp32]
    __shared__ __align__(8) uint64_t a_full[LP_AST], a_empty[LP_AST], w_full, acc_full[2], acc_empty[2], tq_full[LP_TQ],
        tq_empty[LP_TQ];
    __shared__ int tile_q[LP_TQ];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmW); prefetch_tmap(&tmC);
        for (int s = 0; s < LP_AST; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        mbar_init(&w_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        for (int s = 0; s < LP_TQ; ++s) { mbar_init(&tq_full[s], 1); mbar_init(&tq_empty[s], 5); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(&tmem_base_s, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(&w_full, KB * LP_BLK);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * LP_BLK, &tmW, &w_full, kb * 64, 0);
            int it = 0;
            for (int n = 0;; ++n) {
                int tile = a.ticket ? (int)atomicAdd(a.ticket, 1u) : (int)blockIdx.x + n * (int)gridDim.x;
                if (tile >= a.tiles) tile = -1;
                const int qs = n % LP_TQ;
                mbar_wait(&tq_empty[qs], ((n / LP_TQ) & 1) ^ 1);
                tile_q[qs] = tile;
                mbar_arrive(&tq_full[qs]);
                if (tile < 0) break;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % LP_AST;
                    mbar_wait(&a_empty[s], ((it / LP_AST) & 1) ^ 1);
                    mbar_expect_tx(&a_full[s], LP_BLK);
                    tma_load_2d(sA + s * LP_BLK, &tmA, &a_full[s], kb * 64, tile * 128);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_h16(128, LP_N, kF16);
            mbar_wait(&w_full, 0);
            int it = 0;
            for (int n = 0;; ++n) {
                const int qs = n % LP_TQ;
                mbar_wait(&tq_full[qs], (n / LP_TQ) & 1);
                const int tile = tile_q[qs];
                mbar_arrive(&tq_empty[qs]);
                if (tile < 0) break;
                const int acc = n & 1;
                mbar_wait(&acc_empty[acc], ((n >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % LP_AST;
                    mbar_wait(&a_full[s], (it / LP_AST) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(sA + s * LP_BLK), sb = smem_u32(sW + kb * LP_BLK);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16<1>(tmem + acc * LP_N, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sb + kk * 32), idesc,
                                     (kb | kk) ? 1u : 0u);
                    umma_commit(&a_empty[s]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int r_in_tile = q * 32 + lane;
        const bool storer = (warp == 2 && lane == 0);
        int chunk_it = 0;
        for (int n = 0;; ++n) {
            const int qs = n % LP_TQ;
            mbar_wait(&tq_full[qs], (n / LP_TQ) & 1);
            const int tile = tile_q[qs];
            __syncwarp();
            if (lane == 0) mbar_arrive(&tq_empty[qs]);
            if (tile < 0) break;
            const int acc = n & 1;
            const long row = (long)tile * 128 + r_in_tile;
            mbar_wait(&acc_full[acc], (n >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * LP_N;
            float s_sum = 0.f, s_sq = 0.f;
            if constexpr (kBf16Out) {
                // 64-column chunks: [128 rows x 64 bf16] = one 128B-swizzled staging tile; statistics from the fp32 values
#pragma unroll 1
                for (int c0 = 0; c0 < LP_N; c0 += 64, ++chunk_it) {
                    uint32_t pk[32];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        float v[32];
                        tmem_ld32(taddr + c0 + hh * 32, v);
                        if (c0 + 64 == LP_N && hh == 1) {       // accumulator fully read: hand it back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[acc]);
                        }
                        float s = 0.f, qq = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float y0 = v[j] + __ldg(a.bias + c0 + hh * 32 + j), y1 = v[j + 1] + __ldg(a.bias + c0 + hh * 32 + j + 1);
                            s += y0 + y1; qq = fmaf(y0, y0, fmaf(y1, y1, qq));
                            pk[hh * 16 + (j >> 1)] = pack_h16x2<kF16>(y0, y1);
                        }
                        s_sum += s; s_sq += qq;
                    }
                    uint8_t* stage = sC + (chunk_it % LP_CST) * LP_BLK;
                    if (storer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(LP_CST - 1) : "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(stage + sw128_offset(r_in_tile, j)) =
                            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    fence_async_smem();
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (storer) {
                        lp_tma_store_2d(&tmC, stage, c0, tile * 128);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else {
#pragma unroll 1
            for (int c0 = 0; c0 < LP_N; c0 += 32, ++chunk_it) {
                float v[32];
                tmem_ld32(taddr + c0, v);
                if (c0 + 32 == LP_N) {                  // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                }
                float s = 0.f, qq = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] += __ldg(a.bias + c0 + j);
                    s += v[j]; qq = fmaf(v[j], v[j], qq);
                }
                s_sum += s; s_sq += qq;
                uint8_t* stage = sC + (chunk_it % LP_CST) * LP_BLK;
                if (storer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(LP_CST - 1) : "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");          // staging buffer is free again
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stage + sw128_offset(r_in_tile, j)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_async_smem();
                asm volatile("bar.sync 2, 128;" ::: "memory");          // whole [128 x 32] chunk staged
                if (storer) {
                    lp_tma_store_2d(&tmC, stage, c0, tile * 128);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            }
            if (a.stats && row < a.M) a.stats[row] = make_float2(s_sum, s_sq);
        }
        if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, 256);
}

// defined in gemm_tc.cu
int launch_row_stats_finalize(const void* partial, float* mean_rstd, long n_utt, long rows_per_utt, int cols, double eps,
                              cudaStream_t st);

template <int KB, bool kBf16Out, bool kF16>
static int launch_lp(const void* A, const void* W, const float* bias, void* C, int M, void* stats, cudaStream_t st) {
    constexpr int K = KB * 64;
    CUtensorMap tmA, tmW, tmC;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {2, (uint64_t)K * 2};
    const uint32_t bA[2] = {64, 128};
    const uint64_t dW[2] = {(uint64_t)K, (uint64_t)LP_N}, sW[2] = {2, (uint64_t)K * 2};
    const uint32_t bW[2] = {64, (uint32_t)LP_N};
    const uint64_t esz = kBf16Out ? 2 : 4;
    const uint64_t dC[2] = {(uint64_t)LP_N, (uint64_t)M}, sC[2] = {esz, (uint64_t)LP_N * esz};
    const uint32_t bC[2] = {kBf16Out ? 64u : 32u, 128};
    constexpr CUtensorMapDataType t16 = kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (make_tmap(&tmA, t16, 2, A, dA, sA, bA)) return 1;
    if (make_tmap(&tmW, t16, 2, W, dW, sW, bW)) return 1;
    if (make_tmap(&tmC, kBf16Out ? t16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dC, sC, bC)) return 1;
    const size_t smem = (size_t)(KB + LP_AST + LP_CST) * LP_BLK + 1024;
    auto kern = linear_persist_kernel<KB, kBf16Out, kF16>;
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = (int)cdiv(M, 128);
    unsigned* ticket = nullptr;
    if (stats) {        // the ticket lives behind the per-row sums (dprnn_gemm_tc_stats_bytes reserves the room)
        ticket = (unsigned*)((uint8_t*)stats + (size_t)M * sizeof(float2));
        DPRNN_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
    }
    LinPersistArgs args{bias, (float2*)stats, M, tiles, ticket};
    kern<<<tiles < sms ? tiles : sms, 192, smem, st>>>(tmA, tmW, tmC, args);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

}  // namespace dprnn

using namespace dprnn;

static int linear_stats_impl(const void* A, const void* W, const float* bias, void* C, bool bf16_out, int M, int K,
                             void* stats_partial, long rows_per_utt, float eps, float* mean_rstd, void* stream,
                             int h16 = DPRNN_H16_BF16) {
    DPRNN_CHECK_ARG(h16 == DPRNN_H16_BF16 || (h16 == DPRNN_H16_FP16 && bf16_out));
    DPRNN_CHECK_ARG(A && W && bias && C && M > 0 && (K == 128 || K == 256));
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) % 16 == 0);
    if (stats_partial && mean_rstd) DPRNN_CHECK_ARG(rows_per_utt > 0 && M % rows_per_utt == 0);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (h16 == DPRNN_H16_FP16) rc = K == 256 ? launch_lp<4, true, true>(A, W, bias, C, M, stats_partial, st) : launch_lp<2, true, true>(A, W, bias, C, M, stats_partial, st);
    else if (bf16_out) rc = K == 256 ? launch_lp<4, true, false>(A, W, bias, C, M, stats_partial, st) : launch_lp<2, true, false>(A, W, bias, C, M, stats_partial, st);
    else rc = K == 256 ? launch_lp<4, false, false>(A, W, bias, C, M, stats_partial, st) : launch_lp<2, false, false>(A, W, bias, C, M, stats_partial, st);
    if (rc) return rc;
    if (stats_partial && mean_rstd) {     // mean_rstd == NULL: the caller reduces the per-row sums itself (ragged batches)
        return launch_row_stats_finalize(stats_partial, mean_rstd, M / rows_per_utt, rows_per_utt, LP_N, (double)eps, st);
    }
    return 0;
}

extern "C" int dprnn_linear_bf16_stats(const void* A, const void* W, const float* bias, float* C, int M, int K,
                                       void* stats_partial, long rows_per_utt, float eps, float* mean_rstd,
                                       void* stream) {
    return linear_stats_impl(A, W, bias, C, false, M, K, stats_partial, rows_per_utt, eps, mean_rstd, stream);
}

extern "C" int dprnn_linear_bf16out_stats(const void* A, const void* W, const float* bias, void* C_bf16, int M, int K,
                                          void* stats_partial, long rows_per_utt, float eps, float* mean_rstd,
                                          void* stream) {
    return linear_stats_impl(A, W, bias, C_bf16, true, M, K, stats_partial, rows_per_utt, eps, mean_rstd, stream);
}

// dprnn_linear_bf16out_stats with operands and output in the 16-bit format h16 (DPRNN_H16_BF16 / DPRNN_H16_FP16).
extern "C" int dprnn_linear_h16out_stats(const void* A, const void* W, const float* bias, void* C_h16, int M, int K,
                                         void* stats_partial, long rows_per_utt, float eps, float* mean_rstd, int h16,
                                         void* stream) {
    return linear_stats_impl(A, W, bias, C_h16, true, M, K, stats_partial, rows_per_utt, eps, mean_rstd, stream, h16);
}
