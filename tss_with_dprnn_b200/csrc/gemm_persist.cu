// Persistent, pipelined tensor-core contraction for the pointwise (1x1) convolutions of the path with many rows and a
// small weight (bottleneck conv, conv2d after the fold, gated head, end conv, speaker ResNet):
//     C[M, N_out] (fp32) = epi( A[M,K] @ W[N,K]^T + bias )       A, W fp32 read as TF32, or bf16
// These are HBM-bound (K, N <= 256): one CTA per SM keeps W resident in shared memory, streams 128-row tiles of A
// through a TMA ring (tiles handed out by an atomic ticket, so CTAs that start late under multi-stream sharing just
// take fewer), double-buffers the accumulator in TMEM so that the MMA of tile i+1 overlaps the epilogue of tile i, and
// writes C through 128B-swizzled staging + TMA stores (full 128-byte lines instead of per-thread 16-byte fragments).
// The non-persistent gemm_tc_kernel (one CTA per tile, W reloaded per CTA, direct stores) reached 23-52 % of the
// measured HBM bandwidth on these shapes; it remains for the statistics-emitting variant and as the small-M path.
// Warps: 0 = ticket scheduler + TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM lane quadrant = warp % 4).
//
// Operand kinds (argument a_is_bf16 of the C ABI): 0 = fp32 read by the tensor core as TF32, i.e. TRUNCATED to 10
// mantissa bits - a coherent shrink of ~3.5e-4 per operand that measured as the dominant error of the fp16 mode (9.6e-4
// against the reference with TF32 convs, 3.9e-4 with exact ones; tools/accuracy_matrix.py); 1 = bf16;
// 2 = DPRNN_GEMM_F32X2: fp32 operands as PAIRS of bf16 (hi = bf16(a), lo = bf16(a - hi): 16 significand bits) and three
// kind::f16 MMAs per K slice (hi*hi + lo*hi + hi*lo).  A still arrives as fp32 through TMA; warps 6..9 rewrite every
// 128-byte row of a landed K-block IN PLACE as [hi(32) | lo(32)] bf16 (the bytes and the swizzle of the tile do not
// change), W is packed that way once on the host ([N, 2K] bf16).  1.5x the MMA instructions of the TF32 form on a kernel
// that waits for HBM anyway, and the convolutions then sit at 1e-5 instead of 1e-3.
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

constexpr int GP_CST = 2, GP_TQ = 4;
constexpr uint32_t GP_BLK = 128 * 128;          // one [128 rows x 128 B] swizzled tile

struct GemmPersistArgs {
    const float* bias;
    int M, tiles;
    unsigned* ticket;
    long rows_per_utt;         // > 0: bias is per utterance [M / rows_per_utt, N]
    const int* row_utt;        // ragged batches: utterance of every row (per-utterance bias)
    const float *post_scale, *post_shift, *prelu_a;      // DPRNN_EPI_AFFINE_PRELU
};

__device__ __forceinline__ float gp_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gp_sigmoid(float x) { return fmaf(gp_tanh(0.5f * x), 0.5f, 0.5f); }

__host__ __device__ constexpr uint32_t gp_idesc(int elem_bytes, int M, int N) {
    const uint32_t fmt = elem_bytes == 2 ? 1u : 2u;      // BF16 = 1, TF32 = 2
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int kElem>
__device__ __forceinline__ void gp_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
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
__device__ __forceinline__ void gp_tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(m), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
}

// one thread = one row of a [128 x 32 fp32] SWIZZLE_128B K-block: fp32 -> [hi(32) | lo(32)] bf16, in place
__device__ __forceinline__ void gp_split_row(uint8_t* tile, int r) {
    float v[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(tile + sw128_offset(r, c));
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
    }
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        const float2 hf = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
        lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        *reinterpret_cast<uint4*>(tile + sw128_offset(r, c)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        *reinterpret_cast<uint4*>(tile + sw128_offset(r, 4 + c)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
    }
}

// KB = number of 128-byte K-blocks (K * kElem / 128); AST = A ring stages; kSplit: DPRNN_GEMM_F32X2 (kElem == 4)
template <int kElem, int KB, int N, int EPI, int AST, bool kSplit>
__global__ void __launch_bounds__(kSplit ? 320 : 192, 1) gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmW,
                                                              const __grid_constant__ CUtensorMap tmC,
                                                              const GemmPersistArgs a) {
    constexpr int N_OUT = EPI == DPRNN_EPI_GATED ? N / 2 : N;
    constexpr uint32_t W_BLK = N * 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;                                 // KB blocks of [N rows x 128 B]
    uint8_t* sA = sW + KB * W_BLK;                      // AST blocks of [128 rows x 128 B]
    uint8_t* sC = sA + AST * GP_BLK;                    // GP_CST staging blocks [128 rows x 32 fp32]
    __shared__ __align__(8) uint64_t a_full[AST], a_empty[AST], a_conv[AST], w_full, acc_full[2], acc_empty[2],
        tq_full[GP_TQ], tq_empty[GP_TQ];
    __shared__ int tile_q[GP_TQ];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmW); prefetch_tmap(&tmC);
        for (int s = 0; s < AST; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&a_conv[s], 4); }
        mbar_init(&w_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        for (int s = 0; s < GP_TQ; ++s) { mbar_init(&tq_full[s], 1); mbar_init(&tq_empty[s], kSplit ? 9 : 5); }
        fence_barrier_init();
    }
    constexpr uint32_t TMEM_COLS = 2 * N <= 32 ? 32 : 2 * N <= 64 ? 64 : 2 * N <= 128 ? 128 : 2 * N <= 256 ? 256 : 512;
    if (warp == 1) tmem_alloc<1>(&tmem_base_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(&w_full, KB * W_BLK);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * W_BLK, &tmW, &w_full, kb * (kSplit ? 64 : 128 / kElem), 0);
            int it = 0;
            for (int n = 0;; ++n) {
                int tile = (int)atomicAdd(a.ticket, 1u);
                if (tile >= a.tiles) tile = -1;
                const int qs = n % GP_TQ;
                mbar_wait(&tq_empty[qs], ((n / GP_TQ) & 1) ^ 1);
                tile_q[qs] = tile;
                mbar_arrive(&tq_full[qs]);
                if (tile < 0) break;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % AST;
                    mbar_wait(&a_empty[s], ((it / AST) & 1) ^ 1);
                    mbar_expect_tx(&a_full[s], GP_BLK);
                    tma_load_2d(sA + s * GP_BLK, &tmA, &a_full[s], kb * (128 / kElem), tile * 128);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = gp_idesc(kSplit ? 2 : kElem, 128, N);
            mbar_wait(&w_full, 0);
            int it = 0;
            for (int n = 0;; ++n) {
                const int qs = n % GP_TQ;
                mbar_wait(&tq_full[qs], (n / GP_TQ) & 1);
                const int tile = tile_q[qs];
                mbar_arrive(&tq_empty[qs]);
                if (tile < 0) break;
                const int acc = n & 1;
                mbar_wait(&acc_empty[acc], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % AST;
                    mbar_wait(kSplit ? &a_conv[s] : &a_full[s], (it / AST) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(sA + s * GP_BLK), sb = smem_u32(sW + kb * W_BLK);
                    if constexpr (kSplit) {
                        // row = [hi(k..k+31) | lo(k..k+31)] bf16 in both operands: hi*hi + lo*hi + hi*lo per 16-k slice
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const uint64_t ah = umma_desc_sw128(sa + kk * 32), al = umma_desc_sw128(sa + 64 + kk * 32);
                            const uint64_t wh = umma_desc_sw128(sb + kk * 32), wl = umma_desc_sw128(sb + 64 + kk * 32);
                            gp_umma<2>(tmem + acc * N, ah, wh, idesc, (kb | kk) ? 1u : 0u);
                            gp_umma<2>(tmem + acc * N, al, wh, idesc, 1u);
                            gp_umma<2>(tmem + acc * N, ah, wl, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            gp_umma<kElem>(tmem + acc * N, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sb + kk * 32), idesc,
                                           (kb | kk) ? 1u : 0u);
                    }
                    umma_commit(&a_empty[s]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
        __syncwarp();
    } else if (kSplit && warp >= 6) {
        // ================= converter: every landed fp32 K-block of A becomes [hi | lo] bf16 in place =================
        const int r = (warp - 6) * 32 + lane;
        int it = 0;
        for (int n = 0;; ++n) {
            const int qs = n % GP_TQ;
            mbar_wait(&tq_full[qs], (n / GP_TQ) & 1);
            const int tile = tile_q[qs];
            __syncwarp();
            if (lane == 0) mbar_arrive(&tq_empty[qs]);
            if (tile < 0) break;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it % AST;
                mbar_wait(&a_full[s], (it / AST) & 1);
                gp_split_row(sA + s * GP_BLK, r);
                fence_async_smem();              // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_conv[s]);
            }
        }
    } else {
        const int q = warp & 3;
        const int r_in_tile = q * 32 + lane;
        const bool storer = (warp == 2 && lane == 0);
        int chunk_it = 0;
        for (int n = 0;; ++n) {
            const int qs = n % GP_TQ;
            mbar_wait(&tq_full[qs], (n / GP_TQ) & 1);
            const int tile = tile_q[qs];
            __syncwarp();
            if (lane == 0) mbar_arrive(&tq_empty[qs]);
            if (tile < 0) break;
            const int acc = n & 1;
            const long row = (long)tile * 128 + r_in_tile;
            mbar_wait(&acc_full[acc], (n >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * N;
            const float* bias = a.bias;
            if (bias && (a.rows_per_utt > 0 || a.row_utt))
                bias += (row < a.M ? (a.row_utt ? (long)__ldg(a.row_utt + row) : row / a.rows_per_utt) : 0) * (long)N;
#pragma unroll 1
            for (int c0 = 0; c0 < N_OUT; c0 += 32, ++chunk_it) {
                float v[32];
                tmem_ld32(taddr + c0, v);
                if constexpr (EPI == DPRNN_EPI_GATED) {
                    float g[32];
                    tmem_ld32(taddr + N_OUT + c0, g);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        v[j] = gp_tanh(v[j] + __ldg(bias + c0 + j)) * gp_sigmoid(g[j] + __ldg(bias + N_OUT + c0 + j));
                }
                if (c0 + 32 == N_OUT) {                 // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                }
                if constexpr (EPI != DPRNN_EPI_GATED) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = v[j] + (bias ? __ldg(bias + c0 + j) : 0.f);
                        if constexpr (EPI == DPRNN_EPI_RELU) x = fmaxf(x, 0.f);
                        if constexpr (EPI == DPRNN_EPI_SIGMOID) x = gp_sigmoid(x);
                        if constexpr (EPI == DPRNN_EPI_AFFINE_PRELU) {
                            x = fmaf(x, __ldg(a.post_scale + c0 + j), __ldg(a.post_shift + c0 + j));
                            x = x >= 0.f ? x : __ldg(a.prelu_a) * x;
                        }
                        v[j] = x;
                    }
                }
                uint8_t* stage = sC + (chunk_it % GP_CST) * GP_BLK;
                if (storer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(GP_CST - 1) : "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");          // staging buffer is free again
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stage + sw128_offset(r_in_tile, j)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_async_smem();
                asm volatile("bar.sync 2, 128;" ::: "memory");          // whole [128 x 32] chunk staged
                if (storer) {
                    gp_tma_store_2d(&tmC, stage, c0, tile * 128);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, TMEM_COLS);
}

template <int kElem, int KB, int N, int EPI, bool kSplit = false>
static int launch_gp(const void* A, const void* W, float* C, long ldc, const GemmPersistArgs& args0, cudaStream_t st) {
    constexpr int K = KB * 128 / kElem;
    constexpr int N_OUT = EPI == DPRNN_EPI_GATED ? N / 2 : N;
    // shared memory: W (KB x N x 128 B) + A ring + 2 staging tiles + alignment slack <= 227 KB
    constexpr int W_BYTES = KB * N * 128;
    constexpr int AST = (232448 - 1024 - W_BYTES - GP_CST * (int)GP_BLK - 2048) / (int)GP_BLK >= 8 ? 8
                        : (232448 - 1024 - W_BYTES - GP_CST * (int)GP_BLK - 2048) / (int)GP_BLK;
    static_assert(AST >= 2, "weight tile too large for a resident copy");
    const int M = args0.M;
    CUtensorMap tmA, tmW, tmC;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {(uint64_t)kElem, (uint64_t)K * kElem};
    const uint32_t bA[2] = {(uint32_t)(128 / kElem), 128};
    const uint64_t dW[2] = {(uint64_t)K, (uint64_t)N}, sW[2] = {(uint64_t)kElem, (uint64_t)K * kElem};
    const uint32_t bW[2] = {(uint32_t)(128 / kElem), (uint32_t)N};
    const uint64_t dC[2] = {(uint64_t)N_OUT, (uint64_t)M}, sC[2] = {4, (uint64_t)ldc * 4};
    const uint32_t bC[2] = {32, 128};
    const CUtensorMapDataType dt = kElem == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (make_tmap(&tmA, dt, 2, A, dA, sA, bA)) return 1;
    if constexpr (kSplit) {        // W packed on the host as [N, 2K] bf16: per 32 k, [hi(32) | lo(32)]
        const uint64_t dW2[2] = {(uint64_t)2 * K, (uint64_t)N}, sW2[2] = {2, (uint64_t)K * 4};
        const uint32_t bW2[2] = {64, (uint32_t)N};
        if (make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dW2, sW2, bW2)) return 1;
    } else if (make_tmap(&tmW, dt, 2, W, dW, sW, bW)) return 1;
    if (make_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dC, sC, bC)) return 1;
    const size_t smem = (size_t)W_BYTES + (size_t)(AST + GP_CST) * GP_BLK + 1024;
    auto kern = gemm_persist_kernel<kElem, KB, N, EPI, AST, kSplit>;
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    GemmPersistArgs args = args0;
    args.tiles = (int)cdiv(M, 128);
    DPRNN_CUDA(cudaMemsetAsync(args.ticket, 0, sizeof(unsigned), st));
    kern<<<args.tiles < sms ? args.tiles : sms, kSplit ? 320 : 192, smem, st>>>(tmA, tmW, tmC, args);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

template <int kElem, int KB, int N, bool kSplit>
static int gp_dispatch_epi(const void* A, const void* W, float* C, long ldc, const GemmPersistArgs& args, int epi,
                           cudaStream_t st) {
    switch (epi) {
        case DPRNN_EPI_NONE: return launch_gp<kElem, KB, N, DPRNN_EPI_NONE, kSplit>(A, W, C, ldc, args, st);
        case DPRNN_EPI_RELU: return launch_gp<kElem, KB, N, DPRNN_EPI_RELU, kSplit>(A, W, C, ldc, args, st);
        case DPRNN_EPI_SIGMOID: return launch_gp<kElem, KB, N, DPRNN_EPI_SIGMOID, kSplit>(A, W, C, ldc, args, st);
        case DPRNN_EPI_AFFINE_PRELU: return launch_gp<kElem, KB, N, DPRNN_EPI_AFFINE_PRELU, kSplit>(A, W, C, ldc, args, st);
        case DPRNN_EPI_GATED:
            if constexpr (N == 256) return launch_gp<kElem, KB, N, DPRNN_EPI_GATED, kSplit>(A, W, C, ldc, args, st);
            break;
        default: break;
    }
    return -1;
}

template <int kElem, int KB, bool kSplit = false>
static int gp_dispatch_n(const void* A, const void* W, float* C, long ldc, const GemmPersistArgs& args, int N, int epi,
                         cudaStream_t st) {
    if (N == 64) return gp_dispatch_epi<kElem, KB, 64, kSplit>(A, W, C, ldc, args, epi, st);
    if (N == 128) return gp_dispatch_epi<kElem, KB, 128, kSplit>(A, W, C, ldc, args, epi, st);
    if (N == 256) {
        if constexpr (KB <= 4) return gp_dispatch_epi<kElem, KB, 256, kSplit>(A, W, C, ldc, args, epi, st);
    }
    return -1;
}

}  // namespace dprnn

using namespace dprnn;

extern "C" size_t dprnn_gemm_persist_workspace_bytes(void) { return 256; }

// 1 if (elem, N, K, epilogue) is built for the persistent kernel
extern "C" int dprnn_gemm_persist_supported(int a_kind, int N, int K, int epilogue) {
    if (a_kind < 0 || a_kind > DPRNN_GEMM_F32X2) return 0;
    const int a_is_bf16 = a_kind == DPRNN_GEMM_BF16;
    const int kb = K * (a_is_bf16 ? 2 : 4) / 128;
    if ((K * (a_is_bf16 ? 2 : 4)) % 128) return 0;
    if (a_is_bf16 ? !(kb == 1 || kb == 2 || kb == 4) : !(kb == 2 || kb == 4 || kb == 8)) return 0;
    if (!(N == 64 || N == 128 || N == 256)) return 0;
    if (N == 256 && kb > 4) return 0;
    if (epilogue == DPRNN_EPI_GATED && N != 256) return 0;
    return epilogue >= DPRNN_EPI_NONE && epilogue <= DPRNN_EPI_AFFINE_PRELU;
}

extern "C" int dprnn_gemm_persist(const void* A, int a_kind, const void* W, const float* bias, long bias_rows_per_utt,
                                  const int* bias_row_utt, const float* post_scale, const float* post_shift,
                                  const float* prelu_a, float* C, long ldc, int M, int N, int K, int epilogue,
                                  void* workspace, void* stream) {
    DPRNN_CHECK_ARG(A && W && C && workspace && M > 0 && ldc % 4 == 0);
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)W | (uintptr_t)C | (uintptr_t)workspace) % 16 == 0);
    DPRNN_CHECK_ARG(dprnn_gemm_persist_supported(a_kind, N, K, epilogue));
    const int a_is_bf16 = a_kind == DPRNN_GEMM_BF16;
    DPRNN_CHECK_ARG(epilogue != DPRNN_EPI_GATED || bias);
    DPRNN_CHECK_ARG(epilogue != DPRNN_EPI_AFFINE_PRELU || (post_scale && post_shift && prelu_a));
    if (bias_rows_per_utt > 0) DPRNN_CHECK_ARG(bias && M % bias_rows_per_utt == 0 && epilogue != DPRNN_EPI_GATED);
    if (bias_row_utt) DPRNN_CHECK_ARG(bias && epilogue != DPRNN_EPI_GATED);
    GemmPersistArgs args{bias, M, 0, (unsigned*)workspace, bias_rows_per_utt, bias_row_utt, post_scale, post_shift, prelu_a};
    cudaStream_t st = (cudaStream_t)stream;
    const int kb = K * (a_is_bf16 ? 2 : 4) / 128;
    int rc = -1;
    if (a_is_bf16) {
        if (kb == 1) rc = gp_dispatch_n<2, 1>(A, W, C, ldc, args, N, epilogue, st);
        else if (kb == 2) rc = gp_dispatch_n<2, 2>(A, W, C, ldc, args, N, epilogue, st);
        else if (kb == 4) rc = gp_dispatch_n<2, 4>(A, W, C, ldc, args, N, epilogue, st);
    } else if (a_kind == DPRNN_GEMM_F32X2) {
        if (kb == 2) rc = gp_dispatch_n<4, 2, true>(A, W, C, ldc, args, N, epilogue, st);
        else if (kb == 4) rc = gp_dispatch_n<4, 4, true>(A, W, C, ldc, args, N, epilogue, st);
        else if (kb == 8) rc = gp_dispatch_n<4, 8, true>(A, W, C, ldc, args, N, epilogue, st);
    } else {
        if (kb == 2) rc = gp_dispatch_n<4, 2>(A, W, C, ldc, args, N, epilogue, st);
        else if (kb == 4) rc = gp_dispatch_n<4, 4>(A, W, C, ldc, args, N, epilogue, st);
        else if (kb == 8) rc = gp_dispatch_n<4, 8>(A, W, C, ldc, args, N, epilogue, st);
    }
    if (rc == -1) {
        set_error("dprnn_gemm_persist: (kind=%d, N=%d, K=%d, epilogue=%d) is not built", a_kind, N, K, epilogue);
        return 2;
    }
    return rc;
}
