// Deep-K tensor-core contraction of the training step's backward (cfg 5):
//     C[M, 128] (+)= A[M, K] @ W[128, K]^T          A, W fp32 read as TF32 (K % 32 == 0) or bf16 (K % 64 == 0); K = ndir * 4H = 1024
// i.e. d x = d gates @ W_ih, the input gradient of an LSTM layer (backward of src/models/dprnn.py:51-70), accumulated
// straight into the gradient of the residual stream.
//
// The weight (128 x 1024 fp32 = 512 KiB) cannot stay resident as in gemm_persist.cu, and the one-tile-per-CTA kernel
// (gemm_tc.cu) re-reads all of it from L2 for every 128 rows - as many bytes again as the A tile itself; it measured
// 54 % of the HBM bandwidth on this shape.  Here a CTA takes 256 rows at a time, so that every W K-block that streams
// through shared memory serves two M = 128 MMAs (L2 -> SM traffic 1.5x instead of 2x the HBM bytes), runs persistently
// (one CTA per SM, tiles from an atomic ticket), double-buffers the two accumulators in TMEM (4 x 128 columns) so that the
// MMAs of tile i+1 overlap the epilogue of tile i, and writes C through swizzled staging + TMA - with the TMA REDUCE-ADD
// form when accumulating, which removes the separate axpy pass (read C, read the product, write C) and the product's
// round trip through HBM.
// Warps: 0 = ticket scheduler + TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

namespace kd {
constexpr int N = 128, NST = 4, CST = 2, TQ = 4;
constexpr uint32_t BLK = 128 * 128;                  // one [128 rows x 128 B] swizzled block
constexpr uint32_t STAGE = 3 * BLK;                  // A rows 0..127 | A rows 128..255 | W
constexpr uint32_t SMEM = NST * STAGE + CST * BLK + 1024;
static_assert(SMEM <= 232448, "shared memory budget of one SM (227 KiB)");
}  // namespace kd

struct GemmKdeepArgs {
    int M, tiles, num_kb, accumulate;
    unsigned* ticket;
};

template <int kElem>
__device__ __forceinline__ void kd_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if constexpr (kElem == 2) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
}
__device__ __forceinline__ void kd_tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(m), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void kd_tma_reduce_add_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(m), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
}

template <int kElem>      // 4: fp32 operands read as TF32, 2: bf16
__global__ void __launch_bounds__(192, 1) gemm_kdeep_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmW,
                                                            const __grid_constant__ CUtensorMap tmC,
                                                            const GemmKdeepArgs a) {
    using namespace kd;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sC = smem + NST * STAGE;                   // CST staging blocks [128 rows x 32 fp32]
    __shared__ __align__(8) uint64_t full[NST], empty[NST], acc_full[2], acc_empty[2], tq_full[TQ], tq_empty[TQ];
    __shared__ int tile_q[TQ];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmW); prefetch_tmap(&tmC);
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        for (int s = 0; s < TQ; ++s) { mbar_init(&tq_full[s], 1); mbar_init(&tq_empty[s], 5); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int n = 0;; ++n) {
                int tile = (int)atomicAdd(a.ticket, 1u);
                if (tile >= a.tiles) tile = -1;
                const int qs = n % TQ;
                mbar_wait(&tq_empty[qs], ((n / TQ) & 1) ^ 1);
                tile_q[qs] = tile;
                mbar_arrive(&tq_full[qs]);
                if (tile < 0) break;
                for (int kb = 0; kb < a.num_kb; ++kb, ++it) {
                    const int s = it % NST;
                    mbar_wait(&empty[s], ((it / NST) & 1) ^ 1);
                    mbar_expect_tx(&full[s], STAGE);
                    tma_load_2d(smem + s * STAGE, &tmA, &full[s], kb * (128 / kElem), tile * 256);         // [256 rows x 128 B]
                    tma_load_2d(smem + s * STAGE + 2 * BLK, &tmW, &full[s], kb * (128 / kElem), 0);        // [128 rows x 128 B]
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            // fp32 accumulation, both operands K-major, M = 128, N = 128; operand format TF32 = 2, BF16 = 1
            constexpr uint32_t fmt = kElem == 2 ? 1u : 2u;
            constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int it = 0;
            for (int n = 0;; ++n) {
                const int qs = n % TQ;
                mbar_wait(&tq_full[qs], (n / TQ) & 1);
                const int tile = tile_q[qs];
                mbar_arrive(&tq_empty[qs]);
                if (tile < 0) break;
                const int acc = n & 1;
                mbar_wait(&acc_empty[acc], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < a.num_kb; ++kb, ++it) {
                    const int s = it % NST;
                    mbar_wait(&full[s], (it / NST) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE), sb = sa + 2 * BLK;
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            kd_umma<kElem>(tmem + (uint32_t)(acc * 2 + sub) * N, umma_desc_sw128(sa + sub * BLK + kk * 32),
                                         umma_desc_sw128(sb + kk * 32), idesc, (kb | kk) ? 1u : 0u);
                    umma_commit(&empty[s]);
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
            const int qs = n % TQ;
            mbar_wait(&tq_full[qs], (n / TQ) & 1);
            const int tile = tile_q[qs];
            __syncwarp();
            if (lane == 0) mbar_arrive(&tq_empty[qs]);
            if (tile < 0) break;
            const int acc = n & 1;
            mbar_wait(&acc_full[acc], (n >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < 2 * (N / 32); ++ch, ++chunk_it) {
                const int sub = ch / (N / 32), c0 = (ch % (N / 32)) * 32;
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 + sub) * N + c0, v);
                if (ch + 1 == 2 * (N / 32)) {            // both accumulators fully read: hand them back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                }
                uint8_t* stage = sC + (chunk_it % CST) * BLK;
                if (storer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(CST - 1) : "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");          // staging buffer is free again
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stage + sw128_offset(r_in_tile, j)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_async_smem();
                asm volatile("bar.sync 2, 128;" ::: "memory");          // whole [128 x 32] chunk staged
                if (storer) {
                    // rows past M are clipped by the tensor map
                    if (a.accumulate) kd_tma_reduce_add_2d(&tmC, stage, c0, tile * 256 + sub * 128);
                    else kd_tma_store_2d(&tmC, stage, c0, tile * 256 + sub * 128);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

}  // namespace dprnn

using namespace dprnn;

extern "C" size_t dprnn_gemm_kdeep_workspace_bytes(void) { return 256; }

extern "C" int dprnn_gemm_kdeep_supported(int a_is_bf16, int N, int K, long lda, long ldc) {
    const int kblk = a_is_bf16 ? 64 : 32;
    return N == 128 && K >= kblk && K % kblk == 0 && lda % (a_is_bf16 ? 8 : 4) == 0 && lda >= K && ldc % 4 == 0 && ldc >= N;
}

extern "C" int dprnn_gemm_kdeep(const void* A, int a_is_bf16, long lda, const void* W, float* C, long ldc, int M, int N,
                                int K, int accumulate, void* workspace, void* stream) {
    DPRNN_CHECK_ARG(A && W && C && workspace && M > 0 && dprnn_gemm_kdeep_supported(a_is_bf16, N, K, lda, ldc));
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)W | (uintptr_t)C | (uintptr_t)workspace) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t el = a_is_bf16 ? 2 : 4;
    const uint32_t kblk = a_is_bf16 ? 64 : 32;
    const CUtensorMapDataType dt = a_is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUtensorMap tmA, tmW, tmC;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {el, (uint64_t)lda * el};
    const uint32_t bA[2] = {kblk, 256};
    const uint64_t dW[2] = {(uint64_t)K, (uint64_t)N}, sW[2] = {el, (uint64_t)K * el};
    const uint32_t bW[2] = {kblk, (uint32_t)N};
    const uint64_t dC[2] = {(uint64_t)N, (uint64_t)M}, sC[2] = {4, (uint64_t)ldc * 4};
    const uint32_t bC[2] = {32, 128};
    if (make_tmap(&tmA, dt, 2, A, dA, sA, bA)) return 1;
    if (make_tmap(&tmW, dt, 2, W, dW, sW, bW)) return 1;
    if (make_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dC, sC, bC)) return 1;
    auto kern = a_is_bf16 ? gemm_kdeep_kernel<2> : gemm_kdeep_kernel<4>;
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kd::SMEM));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    GemmKdeepArgs args{M, (int)cdiv(M, 256), K / (int)kblk, accumulate ? 1 : 0, (unsigned*)workspace};
    DPRNN_CUDA(cudaMemsetAsync(args.ticket, 0, sizeof(unsigned), st));
    kern<<<args.tiles < sms ? args.tiles : sms, 192, kd::SMEM, st>>>(tmA, tmW, tmC, args);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
