// bf16 tensor-core contraction for the Linear(2H -> F) that follows every LSTM (and any other K-major
// bf16 x bf16 -> fp32 product with N <= 256, K <= 256):  C[M,N] = A[M,K] @ W[N,K]^T + bias.
// A (activations, row-major bf16) and W (nn.Linear layout, bf16) are both K-major, so TMA drops them
// into 128-byte-swizzled shared-memory tiles that tcgen05.mma consumes directly; the fp32 accumulator
// lives in TMEM and is read back with tcgen05.ld for the bias epilogue.
// One CTA per 128-row tile: warp 0 = TMA producer + MMA issuer (one elected thread), warps 0-3 = epilogue.
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

template <int N, int K>
__global__ void __launch_bounds__(128) linear_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                        const __grid_constant__ CUtensorMap tmW,
                                                        const float* __restrict__ bias, float* __restrict__ C, long ldc,
                                                        int M) {
    constexpr int KB = K / 64;                       // 64-element (128-byte) K blocks
    constexpr uint32_t A_BLK = 128 * 128, W_BLK = N * 128;
    constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sW = smem + KB * A_BLK;
    __shared__ __align__(8) uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmW);
        mbar_init(&bar_full, 1);
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(&tmem_base_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(&bar_full, KB * (A_BLK + W_BLK));
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                tma_load_2d(sA + kb * A_BLK, &tmA, &bar_full, kb * 64, m0);
                tma_load_2d(sW + kb * W_BLK, &tmW, &bar_full, kb * 64, 0);
            }
            mbar_wait(&bar_full, 0);
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_bf16(128, N);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {      // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle row
                    const uint64_t da = umma_desc_sw128(smem_u32(sA + kb * A_BLK) + kk * 32);
                    const uint64_t db = umma_desc_sw128(smem_u32(sW + kb * W_BLK) + kk * 32);
                    umma_bf16<1>(tmem, da, db, idesc, (kb | kk) ? 1u : 0u);
                }
            }
            umma_commit(&bar_done);
        }
        __syncwarp();
    }
    mbar_wait(&bar_done, 0);
    tc_fence_after();

    const long row = (long)m0 + warp * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (row < M) {
            float* dst = C + row * ldc + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o;
                o.x = v[j] + (bias ? __ldg(bias + c0 + j) : 0.f);
                o.y = v[j + 1] + (bias ? __ldg(bias + c0 + j + 1) : 0.f);
                o.z = v[j + 2] + (bias ? __ldg(bias + c0 + j + 2) : 0.f);
                o.w = v[j + 3] + (bias ? __ldg(bias + c0 + j + 3) : 0.f);
                *reinterpret_cast<float4*>(dst + j) = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, TMEM_COLS);
}

template <int N, int K>
static int launch_linear(const void* A, const void* W, const float* bias, float* C, long ldc, int M, cudaStream_t st) {
    CUtensorMap tmA, tmW;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {2, (uint64_t)K * 2};
    const uint32_t bA[2] = {64, 128};
    const uint64_t dW[2] = {(uint64_t)K, (uint64_t)N}, sW[2] = {2, (uint64_t)K * 2};
    const uint32_t bW[2] = {64, (uint32_t)N};
    if (make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dA, sA, bA)) return 1;
    if (make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dW, sW, bW)) return 1;
    const size_t smem = (size_t)(K / 64) * (128 * 128 + N * 128) + 1024;
    DPRNN_CUDA(cudaFuncSetAttribute(linear_tc_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linear_tc_kernel<N, K><<<cdiv(M, 128), 128, smem, st>>>(tmA, tmW, bias, C, ldc, M);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

}  // namespace dprnn

using namespace dprnn;

extern "C" int dprnn_linear_bf16(const void* A, const void* W, const float* bias, float* C, long ldc, int M, int N,
                                 int K, void* stream) {
    DPRNN_CHECK_ARG(A && W && C && M > 0 && ldc >= N && ldc % 4 == 0);
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 128 && K == 256) return launch_linear<128, 256>(A, W, bias, C, ldc, M, st);
    if (N == 128 && K == 128) return launch_linear<128, 128>(A, W, bias, C, ldc, M, st);
    if (N == 64 && K == 128) return launch_linear<64, 128>(A, W, bias, C, ldc, M, st);
    if (N == 256 && K == 128) return launch_linear<256, 128>(A, W, bias, C, ldc, M, st);
    set_error("dprnn_linear_bf16: unsupported shape N=%d K=%d", N, K);
    return 2;
}
