// fp32 CUDA-core contraction used for every pointwise (1x1) convolution / Linear of the path in the
// exact-fp32 mode:  C[M,N] = epi( pro(A)[M,K] @ Wt[K,N] + bias ).
//   * A is channels-last activations (row = one frame / one chunk position), Wt the transposed weight.
//   * prologue (optional): a = (a*p_scale[b,k] + p_shift[b,k]) * rowscale[row] + p_add[b,k]
//     - the per-utterance norm + speaker fusion folded into the bottleneck conv.
//   * epilogue: bias per column or per (utterance, column); none / relu / sigmoid / gated tanh*sigmoid.
// 64x128x16 tiles, 256 threads, 4x8 register tile, register-staged double buffering.
#include "common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {

constexpr int BM = 64, BN = 128, BK = 16, AS_LD = BK + 4;

struct GemmParams {
    const float* A; long lda;
    const float* Wt; long ldw;
    float* C; long ldc;
    int M, N, K;
    const float* bias; int bias_per_utt; float bias_scale;
    long rows_per_utt;
    const float* p_scale; const float* p_shift; const float* p_add; const float* rowscale;
    int epi;
    const int* row_utt;      // ragged batches: utterance of every row (else NULL: row / rows_per_utt)
    const float *post_scale, *post_shift, *residual; long ldres;     // DPRNN_EPI_RELU_AFFINE
};

__device__ __forceinline__ float apply_act(float v, int epi) {
    if (epi == DPRNN_EPI_RELU) return fmaxf(v, 0.f);
    if (epi == DPRNN_EPI_SIGMOID) return sigmoid_acc(v);
    return v;
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
    __shared__ __align__(16) float As[2][BM][AS_LD];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // global->register staging assignments
    const int ar = tid >> 2, akq = tid & 3;                 // A: row ar, k = akq*4..+3
    const long arow = m0 + ar;
    const bool arow_ok = arow < p.M;
    const long ab = (p.p_scale || p.p_add) && arow_ok ? (p.row_utt ? (long)__ldg(p.row_utt + arow) : arow / p.rows_per_utt) : 0;
    const float rs = (p.rowscale && arow_ok) ? p.rowscale[arow] : 1.0f;

    float4 areg, breg[2];
    auto load_tiles = [&](int k0) {
        areg = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + akq * 4;
        if (arow_ok && k < p.K) {
            areg = *reinterpret_cast<const float4*>(p.A + arow * p.lda + k);
            if (p.p_scale) {
                const float4 sc = *reinterpret_cast<const float4*>(p.p_scale + ab * p.K + k);
                const float4 sh = *reinterpret_cast<const float4*>(p.p_shift + ab * p.K + k);
                areg.x = fmaf(areg.x, sc.x, sh.x); areg.y = fmaf(areg.y, sc.y, sh.y);
                areg.z = fmaf(areg.z, sc.z, sh.z); areg.w = fmaf(areg.w, sc.w, sh.w);
            }
            if (p.rowscale) { areg.x *= rs; areg.y *= rs; areg.z *= rs; areg.w *= rs; }
            if (p.p_add) {
                const float4 ad = *reinterpret_cast<const float4*>(p.p_add + ab * p.K + k);
                areg.x += ad.x; areg.y += ad.y; areg.z += ad.z; areg.w += ad.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256, br = idx >> 5, bc = (idx & 31) * 4;
            breg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + br < p.K && n0 + bc < p.N)
                breg[i] = __ldg(reinterpret_cast<const float4*>(p.Wt + (long)(k0 + br) * p.ldw + n0 + bc));
        }
    };
    auto store_tiles = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][ar][akq * 4]) = areg;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256, br = idx >> 5, bc = (idx & 31) * 4;
            *reinterpret_cast<float4*>(&Bs[buf][br][bc]) = breg[i];
        }
    };

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (p.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            float4 a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&As[buf][ty * 4 + i][k4 * 4]);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k4 * 4 + kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k4 * 4 + kk][64 + tx * 4]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                    acc[i][0] = fmaf(av, b0.x, acc[i][0]); acc[i][1] = fmaf(av, b0.y, acc[i][1]);
                    acc[i][2] = fmaf(av, b0.z, acc[i][2]); acc[i][3] = fmaf(av, b0.w, acc[i][3]);
                    acc[i][4] = fmaf(av, b1.x, acc[i][4]); acc[i][5] = fmaf(av, b1.y, acc[i][5]);
                    acc[i][6] = fmaf(av, b1.z, acc[i][6]); acc[i][7] = fmaf(av, b1.w, acc[i][7]);
                }
            }
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long row = m0 + ty * 4 + i;
        if (row >= p.M) continue;
        const float* bias = p.bias;
        if (bias && p.bias_per_utt) bias += (p.row_utt ? (long)__ldg(p.row_utt + row) : row / p.rows_per_utt) * p.N;
        if (p.epi == DPRNN_EPI_GATED) {
            const int c0 = n0 + tx * 4, c1 = n0 + 64 + tx * 4;
            if (c1 < p.N) {
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float u = acc[i][j] + (bias ? bias[c0 + j] * p.bias_scale : 0.f);
                    const float g = acc[i][4 + j] + (bias ? bias[c1 + j] * p.bias_scale : 0.f);
                    o[j] = tanhf(u) * sigmoid_acc(g);
                }
                *reinterpret_cast<float4*>(p.C + row * p.ldc + (n0 >> 1) + tx * 4) = make_float4(o[0], o[1], o[2], o[3]);
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = n0 + h * 64 + tx * 4;
                if (c >= p.N) continue;
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float v = acc[i][h * 4 + j] + (bias ? bias[c + j] * p.bias_scale : 0.f);
                    o[j] = apply_act(v, p.epi);
                    if (p.epi == DPRNN_EPI_RELU_AFFINE) {
                        o[j] = fmaxf(v, 0.f);
                        if (p.post_scale) o[j] = fmaf(o[j], p.post_scale[c + j], p.post_shift[c + j]);
                        if (p.residual) o[j] += p.residual[row * p.ldres + c + j];
                    }
                }
                *reinterpret_cast<float4*>(p.C + row * p.ldc + c) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

}  // namespace dprnn

using namespace dprnn;

static int gemm_f32_impl(const float* A, long lda, const float* Wt, long ldw, float* C, long ldc, int M, int N,
                         int K, const float* bias, int bias_per_utt, float bias_scale, long rows_per_utt,
                         const float* p_scale, const float* p_shift, const float* p_add, const float* rowscale,
                         int epilogue, const int* row_utt, void* stream) {
    DPRNN_CHECK_ARG(A && Wt && C && M > 0 && N > 0 && K > 0);
    DPRNN_CHECK_ARG(N % 4 == 0 && K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0);
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)Wt | (uintptr_t)C) % 16 == 0);
    DPRNN_CHECK_ARG(epilogue >= DPRNN_EPI_NONE && epilogue <= DPRNN_EPI_GATED);
    DPRNN_CHECK_ARG(epilogue != DPRNN_EPI_GATED || N % 128 == 0);
    DPRNN_CHECK_ARG((p_scale == nullptr) == (p_shift == nullptr));
    if ((bias_per_utt || p_scale || p_add) && !row_utt) DPRNN_CHECK_ARG(rows_per_utt > 0);
    if (rows_per_utt <= 0) rows_per_utt = M;
    GemmParams p{A, lda, Wt, ldw, C, ldc, M, N, K, bias, bias_per_utt, bias_scale, rows_per_utt,
                 p_scale, p_shift, p_add, rowscale, epilogue, row_utt, nullptr, nullptr, nullptr, 0};
    dim3 grid(cdiv(M, BM), cdiv(N, BN));
    gemm_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_gemm_f32(const float* A, long lda, const float* Wt, long ldw, float* C, long ldc, int M, int N,
                              int K, const float* bias, int bias_per_utt, float bias_scale, long rows_per_utt,
                              const float* p_scale, const float* p_shift, const float* p_add, const float* rowscale,
                              int epilogue, void* stream) {
    return gemm_f32_impl(A, lda, Wt, ldw, C, ldc, M, N, K, bias, bias_per_utt, bias_scale, rows_per_utt, p_scale, p_shift,
                         p_add, rowscale, epilogue, nullptr, stream);
}

extern "C" int dprnn_gemm_f32_ragged(const float* A, long lda, const float* Wt, long ldw, float* C, long ldc, int M,
                                     int N, int K, const float* bias, int bias_per_utt, float bias_scale,
                                     const int* row_utt, const float* p_scale, const float* p_shift, const float* p_add,
                                     const float* rowscale, int epilogue, void* stream) {
    DPRNN_CHECK_ARG(row_utt);
    return gemm_f32_impl(A, lda, Wt, ldw, C, ldc, M, N, K, bias, bias_per_utt, bias_scale, 0, p_scale, p_shift, p_add,
                         rowscale, epilogue, row_utt, stream);
}

// exact-fp32 twin of dprnn_gemm_tc_relu_affine (Wt is the transposed weight [K, N], row stride ldw)
extern "C" int dprnn_gemm_f32_relu_affine(const float* A, long lda, const float* Wt, long ldw, const float* bias,
                                          long bias_rows_per_utt, const float* scale, const float* shift,
                                          const float* residual, long ldres, float* C, long ldc, int M, int N, int K,
                                          void* stream) {
    DPRNN_CHECK_ARG(A && Wt && C && M > 0 && N > 0 && K > 0);
    DPRNN_CHECK_ARG(N % 4 == 0 && K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0);
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)Wt | (uintptr_t)C) % 16 == 0);
    DPRNN_CHECK_ARG((scale == nullptr) == (shift == nullptr));
    DPRNN_CHECK_ARG(bias_rows_per_utt == 0 || (bias && bias_rows_per_utt > 0));
    GemmParams p{A, lda, Wt, ldw, C, ldc, M, N, K, bias, bias_rows_per_utt > 0 ? 1 : 0, 1.0f,
                 bias_rows_per_utt > 0 ? bias_rows_per_utt : (long)M, nullptr, nullptr, nullptr, nullptr,
                 DPRNN_EPI_RELU_AFFINE, nullptr, scale, shift, residual, ldres};
    dim3 grid(cdiv(M, BM), cdiv(N, BN));
    gemm_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
