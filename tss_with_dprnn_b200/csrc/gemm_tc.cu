// Tensor-core contraction for every pointwise (1x1) convolution / Linear of the bf16 mode:
//     C[M,N] = epi( A[M,K] @ W[N,K]^T + bias )          (fp32 accumulation in TMEM)
// A (channels-last activations) and W (nn.Linear / Conv1d layout) are both K-major, so TMA drops 128-byte
// K-blocks of each straight into 128B-swizzled shared-memory tiles that tcgen05.mma consumes.
//   kElem == 2: bf16 operands (kind::f16, 64 elements per K-block)   - the Linear after each LSTM (A = hb)
//   kElem == 4: fp32 operands read as TF32 (kind::tf32, 32 elements per K-block) - head / tail / speaker ResNet,
//               whose inputs live in fp32; no conversion pass is needed.
// One CTA per 128-row tile; K-blocks stream through an NST-stage ring (TMA producer warp -> MMA warp), the
// accumulator is read back with tcgen05.ld by 4 epilogue warps (thread = row).  Epilogues: bias, ReLU, sigmoid,
// gated tanh*sigmoid (N = 2F: columns [out | gate]), and optional per-tile partial sums for the following
// GroupNorm / gLN (so the separate statistics pass over the output disappears).
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

constexpr int GT_NST = 3;

__device__ __forceinline__ float gt_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gt_sigmoid(float x) { return fmaf(gt_tanh(0.5f * x), 0.5f, 0.5f); }

__host__ __device__ constexpr uint32_t umma_idesc(int elem_bytes, int M, int N) {
    const uint32_t fmt = elem_bytes == 2 ? 1u : 2u;      // BF16 = 1, TF32 = 2
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int kElem>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if constexpr (kElem == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
}

struct GemmTcArgs {
    const float* bias;
    float* C; long ldc;
    int M; int num_kb;
    float2* stats_partial;     // [M] per-row {sum, sumsq} over the N output columns, or NULL
    long rows_per_utt;
    int bias_per_utt;          // bias is [M / rows_per_utt, N] (a per-utterance bias: the 'cat' speaker fusion, A.6)
    const int* row_utt;        // ragged batches: utterance of every row (else NULL: row / rows_per_utt)
    const float *post_scale, *post_shift, *prelu_a;      // DPRNN_EPI_AFFINE_PRELU / DPRNN_EPI_RELU_AFFINE
    const float* residual; long ldres;                    // DPRNN_EPI_RELU_AFFINE: added after the affine
};

template <int kElem, int N, int EPI>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmW, const GemmTcArgs a) {
    constexpr uint32_t A_BLK = 128 * 128, W_BLK = N * 128, STAGE = A_BLK + W_BLK;
    constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    constexpr int N_OUT = EPI == DPRNN_EPI_GATED ? N / 2 : N;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[GT_NST], bar_empty[GT_NST], bar_done;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmW);
        for (int s = 0; s < GT_NST; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
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
            for (int kb = 0; kb < a.num_kb; ++kb) {
                const int s = kb % GT_NST;
                mbar_wait(&bar_empty[s], ((kb / GT_NST) & 1) ^ 1);
                mbar_expect_tx(&bar_full[s], STAGE);
                tma_load_2d(smem + s * STAGE, &tmA, &bar_full[s], kb * (128 / kElem), m0);
                tma_load_2d(smem + s * STAGE + A_BLK, &tmW, &bar_full[s], kb * (128 / kElem), 0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(kElem, 128, N);
            for (int kb = 0; kb < a.num_kb; ++kb) {
                const int s = kb % GT_NST;
                mbar_wait(&bar_full[s], (kb / GT_NST) & 1);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE), sb = sa + A_BLK;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_ss<kElem>(tmem, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sb + kk * 32), idesc,
                                   (kb | kk) ? 1u : 0u);
                umma_commit(&bar_empty[s]);
            }
            umma_commit(&bar_done);
        }
        __syncwarp();
    } else {
        // epilogue warps 2..5 -> TMEM lane quadrants (warp % 4)
        const int q = warp & 3;
        const long row = (long)m0 + q * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        mbar_wait(&bar_done, 0);
        tc_fence_after();
        float s_sum = 0.f, s_sq = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < N_OUT; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + c0, v);
            if constexpr (EPI == DPRNN_EPI_GATED) {
                float g[32];
                tmem_ld32(taddr + N_OUT + c0, g);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = gt_tanh(v[j] + __ldg(a.bias + c0 + j)) * gt_sigmoid(g[j] + __ldg(a.bias + N_OUT + c0 + j));
            } else {
                const float* bias = a.bias;
                if (bias && a.bias_per_utt)
                    bias += (row < a.M ? (a.row_utt ? (long)__ldg(a.row_utt + row) : row / a.rows_per_utt) : 0) * (long)N;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = v[j] + (bias ? __ldg(bias + c0 + j) : 0.f);
                    if constexpr (EPI == DPRNN_EPI_RELU) x = fmaxf(x, 0.f);
                    if constexpr (EPI == DPRNN_EPI_SIGMOID) x = gt_sigmoid(x);
                    if constexpr (EPI == DPRNN_EPI_AFFINE_PRELU) {
                        x = fmaf(x, __ldg(a.post_scale + c0 + j), __ldg(a.post_shift + c0 + j));
                        x = x >= 0.f ? x : __ldg(a.prelu_a) * x;
                    }
                    if constexpr (EPI == DPRNN_EPI_RELU_AFFINE) {
                        x = fmaxf(x, 0.f);
                        if (a.post_scale) x = fmaf(x, __ldg(a.post_scale + c0 + j), __ldg(a.post_shift + c0 + j));
                        if (a.residual && row < a.M) x += __ldg(a.residual + row * a.ldres + c0 + j);
                    }
                    v[j] = x;
                }
            }
            if (row < a.M) {
                float* dst = a.C + row * a.ldc + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (a.stats_partial) {
                    float s = 0.f, qq = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) { s += v[j]; qq = fmaf(v[j], v[j], qq); }
                    s_sum += s; s_sq += qq;
                }
            }
        }
        // per-row partials: independent of how rows fall into tiles, so a batched call reduces exactly like a
        // per-utterance call (bit-identical results however the batch is sharded)
        if (a.stats_partial && row < a.M) a.stats_partial[row] = make_float2(s_sum, s_sq);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, TMEM_COLS);
}

// mean / rstd per utterance from the per-row partials written by the GEMM epilogue (fixed reduction order)
__global__ void row_stats_finalize_kernel(const float2* __restrict__ partial, float* __restrict__ mean_rstd,
                                          long rows_per_utt, int cols, double eps) {
    __shared__ double scratch[32];
    const int b = blockIdx.x;
    const float2* p = partial + (long)b * rows_per_utt;
    double s = 0.0, q = 0.0;
    for (long r = threadIdx.x; r < rows_per_utt; r += 256) {
        const float2 v = p[r];
        s += (double)v.x; q += (double)v.y;
    }
    s = block_sum(s, scratch);
    q = block_sum(q, scratch);
    if (threadIdx.x == 0) {
        const double cnt = (double)rows_per_utt * cols;
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        mean_rstd[2 * b] = (float)mean;
        mean_rstd[2 * b + 1] = (float)(1.0 / sqrt(var + eps));
    }
}

// host-side launcher shared with linear_persist.cu
int launch_row_stats_finalize(const void* partial, float* mean_rstd, long n_utt, long rows_per_utt, int cols, double eps,
                              cudaStream_t st) {
    row_stats_finalize_kernel<<<(unsigned)n_utt, 256, 0, st>>>((const float2*)partial, mean_rstd, rows_per_utt, cols, eps);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

template <int kElem, int N, int EPI>
static int launch_gemm_tc(const void* A, const void* W, const GemmTcArgs& args, int K, cudaStream_t st) {
    CUtensorMap tmA, tmW;
    const uint32_t kblk = 128 / kElem;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)args.M}, sA[2] = {(uint64_t)kElem, (uint64_t)K * kElem};
    const uint32_t bA[2] = {kblk, 128};
    const uint64_t dW[2] = {(uint64_t)K, (uint64_t)N}, sW[2] = {(uint64_t)kElem, (uint64_t)K * kElem};
    const uint32_t bW[2] = {kblk, (uint32_t)N};
    const CUtensorMapDataType dt = kElem == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (make_tmap(&tmA, dt, 2, A, dA, sA, bA)) return 1;
    if (make_tmap(&tmW, dt, 2, W, dW, sW, bW)) return 1;
    const size_t smem = (size_t)GT_NST * (128 * 128 + N * 128) + 1024;
    auto kern = gemm_tc_kernel<kElem, N, EPI>;
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<cdiv(args.M, 128), 192, smem, st>>>(tmA, tmW, args);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

template <int kElem, int N>
static int dispatch_epi(const void* A, const void* W, const GemmTcArgs& args, int K, int epi, cudaStream_t st) {
    switch (epi) {
        case DPRNN_EPI_NONE: return launch_gemm_tc<kElem, N, DPRNN_EPI_NONE>(A, W, args, K, st);
        case DPRNN_EPI_RELU: return launch_gemm_tc<kElem, N, DPRNN_EPI_RELU>(A, W, args, K, st);
        case DPRNN_EPI_SIGMOID: return launch_gemm_tc<kElem, N, DPRNN_EPI_SIGMOID>(A, W, args, K, st);
        case DPRNN_EPI_AFFINE_PRELU:
            if constexpr (kElem == 4 && N >= 128) return launch_gemm_tc<kElem, N, DPRNN_EPI_AFFINE_PRELU>(A, W, args, K, st);
            break;
        case DPRNN_EPI_RELU_AFFINE:
            if constexpr (kElem == 4 && N >= 128) return launch_gemm_tc<kElem, N, DPRNN_EPI_RELU_AFFINE>(A, W, args, K, st);
            break;
        default: break;
    }
    set_error("dprnn_gemm_tc: epilogue %d not built for N=%d", epi, N);
    return 2;
}

}  // namespace dprnn

using namespace dprnn;

extern "C" size_t dprnn_gemm_tc_stats_bytes(int M) { return (size_t)M * sizeof(float2) + 256; }   // + scheduler ticket

static int gemm_tc_impl(const void* A, int a_is_bf16, const void* W, const float* bias, float* C, long ldc, int M,
                        int N, int K, int epilogue, void* stats_partial, long rows_per_utt, float eps,
                        float* mean_rstd, const int* row_utt, void* stream) {
    DPRNN_CHECK_ARG(A && W && C && M > 0 && N > 0 && K > 0 && ldc % 4 == 0);
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) % 16 == 0);
    const int elem = a_is_bf16 ? 2 : 4;
    DPRNN_CHECK_ARG((K * elem) % 128 == 0);
    DPRNN_CHECK_ARG(epilogue != DPRNN_EPI_GATED || bias);
    cudaStream_t st = (cudaStream_t)stream;
    // without stats_partial, rows_per_utt > 0 selects a per-utterance bias [M / rows_per_utt, N]
    const int bias_per_utt = (!stats_partial && (rows_per_utt > 0 || row_utt) && bias) ? 1 : 0;
    if (bias_per_utt) DPRNN_CHECK_ARG((row_utt || M % rows_per_utt == 0) && epilogue != DPRNN_EPI_GATED);
    GemmTcArgs args{bias, C, ldc, M, K * elem / 128, (float2*)stats_partial, rows_per_utt, bias_per_utt, row_utt,
                    nullptr, nullptr, nullptr, nullptr, 0};
    if (stats_partial) {
        DPRNN_CHECK_ARG(rows_per_utt > 0 && M % rows_per_utt == 0 && mean_rstd && epilogue == DPRNN_EPI_NONE);
    }
    int rc = 2;
    if (epilogue == DPRNN_EPI_GATED) {
        if (N == 256 && elem == 4) rc = launch_gemm_tc<4, 256, DPRNN_EPI_GATED>(A, W, args, K, st);
        else set_error("dprnn_gemm_tc: gated epilogue is built for N=256 fp32 operands");
    } else if (elem == 2) {
        if (N == 128) rc = dispatch_epi<2, 128>(A, W, args, K, epilogue, st);
        else set_error("dprnn_gemm_tc: bf16 operands are built for N=128 (got %d)", N);
    } else {
        if (N == 64) rc = dispatch_epi<4, 64>(A, W, args, K, epilogue, st);
        else if (N == 128) rc = dispatch_epi<4, 128>(A, W, args, K, epilogue, st);
        else if (N == 256) rc = dispatch_epi<4, 256>(A, W, args, K, epilogue, st);
        else set_error("dprnn_gemm_tc: N must be 64, 128 or 256 (got %d)", N);
    }
    if (rc) return rc;
    if (stats_partial) {
        return launch_row_stats_finalize(stats_partial, mean_rstd, M / rows_per_utt, rows_per_utt, N, (double)eps, st);
    }
    return 0;
}

extern "C" int dprnn_gemm_tc(const void* A, int a_is_bf16, const void* W, const float* bias, float* C, long ldc, int M,
                             int N, int K, int epilogue, void* stats_partial, long rows_per_utt, float eps,
                             float* mean_rstd, void* stream) {
    return gemm_tc_impl(A, a_is_bf16, W, bias, C, ldc, M, N, K, epilogue, stats_partial, rows_per_utt, eps, mean_rstd,
                        nullptr, stream);
}

// per-utterance bias [n_utt, N] selected through row_utt (ragged batches)
extern "C" int dprnn_gemm_tc_ragged(const void* A, int a_is_bf16, const void* W, const float* bias_per_utt,
                                    const int* row_utt, float* C, long ldc, int M, int N, int K, int epilogue,
                                    void* stream) {
    DPRNN_CHECK_ARG(bias_per_utt && row_utt);
    return gemm_tc_impl(A, a_is_bf16, W, bias_per_utt, C, ldc, M, N, K, epilogue, nullptr, 0, 0.f, nullptr, row_utt, stream);
}

extern "C" int dprnn_gemm_tc_affine_prelu(const void* A, const void* W, const float* scale, const float* shift,
                                          const float* prelu_a, float* C, long ldc, int M, int N, int K, void* stream) {
    DPRNN_CHECK_ARG(A && W && scale && shift && prelu_a && C && M > 0 && (N == 128 || N == 256) && K > 0 && ldc % 4 == 0);
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) % 16 == 0 && (K * 4) % 128 == 0);
    GemmTcArgs args{nullptr, C, ldc, M, K * 4 / 128, nullptr, 0, 0, nullptr, scale, shift, prelu_a, nullptr, 0};
    return N == 128 ? dispatch_epi<4, 128>(A, W, args, K, DPRNN_EPI_AFFINE_PRELU, (cudaStream_t)stream)
                    : dispatch_epi<4, 256>(A, W, args, K, DPRNN_EPI_AFFINE_PRELU, (cudaStream_t)stream);
}

extern "C" int dprnn_gemm_tc_relu_affine(const void* A, const void* W, const float* bias, long bias_rows_per_utt,
                                         const float* scale, const float* shift, const float* residual, long ldres,
                                         float* C, long ldc, int M, int N, int K, void* stream) {
    DPRNN_CHECK_ARG(A && W && C && M > 0 && (N == 128 || N == 256) && K > 0 && ldc % 4 == 0);
    DPRNN_CHECK_ARG(bias_rows_per_utt == 0 || (bias && bias_rows_per_utt > 0 && M % bias_rows_per_utt == 0));
    DPRNN_CHECK_ARG((scale == nullptr) == (shift == nullptr));
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)W | (uintptr_t)C) % 16 == 0 && (K * 4) % 128 == 0);
    GemmTcArgs args{bias, C, ldc, M, K * 4 / 128, nullptr, bias_rows_per_utt, bias_rows_per_utt > 0 ? 1 : 0, nullptr,
                    scale, shift, nullptr, residual, ldres};
    return N == 128 ? dispatch_epi<4, 128>(A, W, args, K, DPRNN_EPI_RELU_AFFINE, (cudaStream_t)stream)
                    : dispatch_epi<4, 256>(A, W, args, K, DPRNN_EPI_RELU_AFFINE, (cudaStream_t)stream);
}
