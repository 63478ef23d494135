// Weight-gradient contraction of the training step (cfg 5) on the tensor cores:
//     C[N1,N2] (+)= A[M,N1]^T B[M,N2]        (M = every chunk position of the batch, ~10^6; N1, N2 <= 1024)
// e.g. dW_ih = dgates^T x, dW_hh = dgates^T h_{t-1}, dW_linear = dy^T h  (backward of src/models/dprnn.py:51-70).
// Both operands are row-major activations, i.e. the contraction index (the row) is the SLOW index: for tcgen05 they
// are MN-major operands.  For 32-bit MN-major operands the only swizzled shared-memory layout the tensor core reads is
// SWIZZLE_128B with a 32-byte base (32-byte chunks of each 128-byte line XOR-ed with the line index mod 4).  TMA writes
// exactly that with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: [32 rows x 32 columns] fp32 boxes land as 128-byte lines,
// 4-line groups of 512 B (= SBO), column groups of 32 one box apart (= LBO); tcgen05.mma kind::tf32 consumes them with
// the a_major / b_major bits set - no transposition pass, no conversion pass.
// One operand (128 columns wide) takes the MMA's M role, the other is cut into N tiles of 256 (or 128) columns.  grid = (column tiles, row splits) ~ one CTA per SM; every CTA streams its row range through
// a 4-stage TMA ring, accumulates in TMEM (fp32) and writes one [tile, 128] partial; a second kernel adds the
// partials in a fixed order (deterministic) into C.
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

constexpr int AB_KB = 32;        // contraction rows per stage (4 MMAs of K = 8)
constexpr int AB_NST = 4;
constexpr uint32_t AB_GROUP = AB_KB * 128;      // bytes of one 32-column group of a stage = LBO

// MN-major SWIZZLE_128B_BASE32B operand: start>>4 [0,14), LBO>>4 [16,30) = distance between 32-column groups, SBO>>4
// [32,46) = distance between 4-row groups (512), version 1 [46,48), layout SWIZZLE_128B_BASE32B = 1 [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(AB_GROUP >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// kind::tf32, D = f32, A and B MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(192) atb_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                     const __grid_constant__ CUtensorMap tmY, long M,
                                                     long rows_per_split, int nyt, int ycols,
                                                     float* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[AB_NST], bar_empty[AB_NST], bar_done;
    __shared__ uint32_t tmem_base_s;
    const uint32_t X_BYTES = 4 * AB_GROUP, Y_BYTES = (uint32_t)(nyt / 32) * AB_GROUP, STAGE = X_BYTES + Y_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    const long r0 = (long)split * rows_per_split;
    const long r1 = r0 + rows_per_split < M ? r0 + rows_per_split : M;
    const int num_kb = r1 > r0 ? (int)((r1 - r0 + AB_KB - 1) / AB_KB) : 0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmY);
        for (int s = 0; s < AB_NST; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(&tmem_base_s, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % AB_NST;
                mbar_wait(&bar_empty[s], ((kb / AB_NST) & 1) ^ 1);
                mbar_expect_tx(&bar_full[s], STAGE);
                const int row = (int)(r0 + (long)kb * AB_KB);
                tma_load_3d(smem + s * STAGE, &tmX, &bar_full[s], 0, row, 0);
                tma_load_3d(smem + s * STAGE + X_BYTES, &tmY, &bar_full[s], 0, row, tile * (nyt / 32));
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_tf32_mn(128, nyt);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % AB_NST;
                mbar_wait(&bar_full[s], (kb / AB_NST) & 1);
                tc_fence_after();
                const uint32_t sx = smem_u32(smem + s * STAGE), sy = sx + X_BYTES;
#pragma unroll
                for (int kk = 0; kk < AB_KB / 8; ++kk) {
                    const uint32_t acc = (kb | kk) ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem), "l"(umma_desc_sw128_mn(sx + kk * 1024)), "l"(umma_desc_sw128_mn(sy + kk * 1024)),
                          "r"(idesc), "r"(acc) : "memory");
                }
                umma_commit(&bar_empty[s]);
            }
            umma_commit(&bar_done);
        }
        __syncwarp();
    } else {
        // epilogue warps 2..5 -> TMEM lane quadrant warp % 4; lane = column of X, TMEM column = column of Y
        const int q = warp & 3;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        float* dst = partial + ((long)split * ycols + (long)tile * nyt) * 128 + q * 32 + lane;
        if (num_kb > 0) {
            mbar_wait(&bar_done, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int c0 = 0; c0 < nyt; c0 += 32) {
            float v[32];
            if (num_kb > 0) {
                tmem_ld32(taddr + c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[(long)(c0 + j) * 128] = v[j];       // 32 lanes -> 128 contiguous bytes
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, 256);
}

// C (+)= sum over splits (fixed order).  partial[split][y][x]; x_is_row: C[x, y] else C[y, x]
__global__ void atb_tc_reduce_kernel(const float* __restrict__ partial, int splits, int ycols, float* __restrict__ C,
                                     long ldc, int x_is_row, int accumulate) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= (long)ycols * 128) return;
    const int y = (int)(e >> 7), x = (int)(e & 127);
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(long)k * ycols * 128 + e];
    float* dst = x_is_row ? C + (long)x * ldc + y : C + (long)y * ldc + x;
    *dst = accumulate ? *dst + s : s;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same contraction with the COLUMN SUMS of A for free:  C[N1,128] (+)= A[M,N1]^T B[M,128],  colsum[N1] (+)= sum_m A[m,:]
// (dW_ih = dgates^T x together with db = sum dgates, backward of src/models/dprnn.py:23-28).  Here a 128-column tile of A
// takes the MMA's M role and B the N role, extended by one 32-column group of ONES that lives in shared memory (written
// once per stage slot, never touched by TMA): D[a-column, 128 + 0] = sum_m A[m, a-column] * 1.  The bias gradient then costs
// no pass of its own over the 3.2 GB d-gates tensor (it was a separate 0.6 ms reduction per half-block).
// grid = (N1 / 128 tiles, row splits); partial[split][tile][160][128].
constexpr int AB_YG = 5;         // B's four 32-column groups + the ones group

__global__ void __launch_bounds__(192) atb_tc_colsum_kernel(const __grid_constant__ CUtensorMap tmX,
                                                            const __grid_constant__ CUtensorMap tmY, long M,
                                                            long rows_per_split, int tiles, float* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[AB_NST], bar_empty[AB_NST], bar_done;
    __shared__ uint32_t tmem_base_s;
    constexpr uint32_t X_BYTES = 4 * AB_GROUP, Y_BYTES = AB_YG * AB_GROUP, STAGE = X_BYTES + Y_BYTES, Y_TMA = 4 * AB_GROUP;
    constexpr int NY = 32 * AB_YG;                     // 160 accumulator columns

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    const long r0 = (long)split * rows_per_split;
    const long r1 = r0 + rows_per_split < M ? r0 + rows_per_split : M;
    const int num_kb = r1 > r0 ? (int)((r1 - r0 + AB_KB - 1) / AB_KB) : 0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmY);
        for (int s = 0; s < AB_NST; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    // the ones group of every stage slot (rows past M count too: TMA zero-fills A's out-of-range rows, 0 * 1 = 0)
    for (int i = threadIdx.x; i < AB_NST * (int)(AB_GROUP / 4); i += blockDim.x) {
        const int s = i / (int)(AB_GROUP / 4), j = i % (int)(AB_GROUP / 4);
        reinterpret_cast<float*>(smem + s * STAGE + X_BYTES + Y_TMA)[j] = 1.0f;
    }
    fence_async_smem();
    if (warp == 1) tmem_alloc<1>(&tmem_base_s, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % AB_NST;
                mbar_wait(&bar_empty[s], ((kb / AB_NST) & 1) ^ 1);
                mbar_expect_tx(&bar_full[s], X_BYTES + Y_TMA);
                const int row = (int)(r0 + (long)kb * AB_KB);
                tma_load_3d(smem + s * STAGE, &tmX, &bar_full[s], 0, row, tile * 4);
                tma_load_3d(smem + s * STAGE + X_BYTES, &tmY, &bar_full[s], 0, row, 0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_tf32_mn(128, NY);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % AB_NST;
                mbar_wait(&bar_full[s], (kb / AB_NST) & 1);
                tc_fence_after();
                const uint32_t sx = smem_u32(smem + s * STAGE), sy = sx + X_BYTES;
#pragma unroll
                for (int kk = 0; kk < AB_KB / 8; ++kk) {
                    const uint32_t acc = (kb | kk) ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem), "l"(umma_desc_sw128_mn(sx + kk * 1024)), "l"(umma_desc_sw128_mn(sy + kk * 1024)),
                          "r"(idesc), "r"(acc) : "memory");
                }
                umma_commit(&bar_empty[s]);
            }
            umma_commit(&bar_done);
        }
        __syncwarp();
    } else {
        // epilogue warps 2..5 -> TMEM lane quadrant warp % 4; lane = column of the A tile, TMEM column = column of [B | 1]
        const int q = warp & 3;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        float* dst = partial + (((long)split * tiles + tile) * NY) * 128 + q * 32 + lane;
        if (num_kb > 0) {
            mbar_wait(&bar_done, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int c0 = 0; c0 < NY; c0 += 32) {
            float v[32];
            if (num_kb > 0) {
                tmem_ld32(taddr + c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (c0 < 128) {
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[(long)(c0 + j) * 128] = v[j];
            } else {
                dst[(long)128 * 128] = v[0];            // the ones group: 32 identical columns, one is enough
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, 256);
}

// C[a, b] (+)= sum over splits, colsum[a] (+)= likewise (fixed order: deterministic)
__global__ void atb_tc_colsum_reduce_kernel(const float* __restrict__ partial, int splits, int tiles, float* __restrict__ C,
                                            long ldc, float* __restrict__ colsum, int accumulate, int accumulate_colsum) {
    constexpr int NY = 32 * AB_YG;
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long per_tile = 129L * 128;                   // 128 columns of B + the column-sum row
    if (e >= (long)tiles * per_tile) return;
    const int tile = (int)(e / per_tile), r = (int)(e % per_tile), y = r >> 7, x = r & 127;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(((long)k * tiles + tile) * NY + y) * 128 + x];
    const long a = (long)tile * 128 + x;
    if (y < 128) {
        float* dst = C + a * ldc + y;
        *dst = accumulate ? *dst + s : s;
    } else {
        colsum[a] = accumulate_colsum ? colsum[a] + s : s;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Both weight gradients of an LSTM direction and its bias gradient from ONE pass over d gates:
//     C1[N1,128] (+)= A^T B1,   C2[N1,128] (+)= A^T shift_t(B2),   colsum[N1] (+)= sum_rows A
// with A = d gates of the direction (N1 = 512), B1 = x (dW_ih), B2 = the layer's fp32 output h read ONE TIME STEP EARLIER
// (dW_hh = sum_t dgates_t^T h_{t-1}; the reverse direction reads one step LATER) - backward of src/models/dprnn.py:23-28.
// The training step is bound by its aggregate HBM traffic (the weight-gradient passes run on a side stream next to the
// BPTT chain: taking them out shortens the step from 107 to 81 ms), and d gates is its largest tensor: 3.2 GB per
// half-block at 16 utterances.  The separate passes read it twice (dW_ih + bias, dW_hh) and needed a shifted COPY of h
// (dprnn_shift_rows: 0.8 GB read + 0.8 GB written).  Here the rows are addressed as (time, sequence) through 5-D tensor
// maps {32 columns, d1, d2, d3, column groups} - intra-chunk layer: d1 = time, d2 = sequence; inter-chunk: d1 = position in
// the chunk, d2 = time, d3 = utterance - and a contraction block is 32 consecutive d1 indices of one (d2, d3).  B2's box is
// fetched at time coordinate t -/+ 1: the step before the first (after the last) is out of bounds, which TMA fills with
// zeros - exactly h_{-1} = 0.  The same kernel with shift 0 and B1 | B2 = the two halves of h gives the Linear's
// dW = dy^T h and db = sum dy in one pass over dy.
// B operand in shared memory = [B1 (4 groups) | B2 (4 groups) | ones (1 group)]: one N = 256 and one N = 32 MMA per K slice.
// kElem = 2: bf16 operands (d gates as dprnn_lstm_bptt_tc_bf16out writes it, x and h as the bf16 copies the tensor-core
// forward works on) - half the bytes of the fp32 form.  16-bit MN-major operands use the plain SWIZZLE_128B layout: a group
// is 64 columns (128 bytes) wide, 8 contraction rows form the 1024-byte swizzle atom, an MMA consumes K = 16 rows.
constexpr int AD_ROWS = 257;     // rows of a partial: 256 columns of [B1 | B2] + the column-sum row

struct AtbDualGeo {
    int n1chunks, D2;            // blocks of KB indices along d1; extent of d2
    int sh1, sh2;                // B2's coordinate offset along d1 / d2 (the time shift)
    long total_chunks, chunks_per_split;
};

__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int kElem>
struct AdCfg {
    static constexpr int KB = kElem == 2 ? 64 : 32;                 // contraction rows per stage
    static constexpr int G128 = kElem == 2 ? 2 : 4;                 // column groups (128 bytes wide) per 128 columns
    static constexpr uint32_t GROUP = KB * 128;                     // bytes of one column group of a stage = LBO
    static constexpr uint32_t STAGE = (3 * G128 + 1) * GROUP;       // A | B1 | B2 | ones
    static constexpr uint32_t TMA_BYTES = 3 * G128 * GROUP;
    static constexpr uint32_t KSTEP = kElem == 2 ? 2048 : 1024;     // bytes of the rows one MMA consumes (K = 16 / 8)
    static constexpr size_t SMEM = (size_t)AB_NST * STAGE + 1024;
};
static_assert(AdCfg<2>::SMEM <= 232448 && AdCfg<4>::SMEM <= 232448, "shared memory budget of one SM (227 KiB)");

template <int kElem>
__device__ __forceinline__ uint64_t ad_desc_mn(uint32_t smem_addr) {
    if constexpr (kElem == 2)       // SWIZZLE_128B (2): LBO = distance between 64-column groups, SBO = between 8-row atoms
        return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(AdCfg<2>::GROUP >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
               ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    else
        return umma_desc_sw128_mn(smem_addr);
}
template <int kElem>
__device__ __forceinline__ void ad_umma(uint32_t tmem_d, uint64_t da, uint64_t db, int n, uint32_t acc) {
    if constexpr (kElem == 2) {
        // kind::f16, D = f32, A = B = bf16, both MN-major (bits 15, 16)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    } else {
        const uint32_t idesc = umma_idesc_tf32_mn(128, n);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
}

template <int kElem>
__global__ void __launch_bounds__(192) atb_dual_kernel(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB1,
                                                       const __grid_constant__ CUtensorMap tmB2, const AtbDualGeo g,
                                                       int tiles, float* __restrict__ partial) {
    using Cf = AdCfg<kElem>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[AB_NST], bar_empty[AB_NST], bar_done;
    __shared__ uint32_t tmem_base_s;
    constexpr uint32_t X_BYTES = Cf::G128 * Cf::GROUP, STAGE = Cf::STAGE, TMA_BYTES = Cf::TMA_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    const long k0 = (long)split * g.chunks_per_split;
    const long k1 = k0 + g.chunks_per_split < g.total_chunks ? k0 + g.chunks_per_split : g.total_chunks;
    const int num_kb = k1 > k0 ? (int)(k1 - k0) : 0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB1); prefetch_tmap(&tmB2);
        for (int s = 0; s < AB_NST; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    // the ones group of every stage slot (A's out-of-range rows are zero-filled by TMA: 0 * 1 = 0)
    for (int i = threadIdx.x; i < AB_NST * (int)(Cf::GROUP / 4); i += blockDim.x) {
        const int s = i / (int)(Cf::GROUP / 4), j = i % (int)(Cf::GROUP / 4);
        reinterpret_cast<uint32_t*>(smem + s * STAGE + TMA_BYTES)[j] = kElem == 2 ? 0x3F803F80u : 0x3F800000u;
    }
    fence_async_smem();
    if (warp == 1) tmem_alloc<1>(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            // chunk -> (d1 block, d2, d3), kept incrementally
            long rest = k0 / g.n1chunks;
            int i1 = (int)(k0 % g.n1chunks), i2 = (int)(rest % g.D2), i3 = (int)(rest / g.D2);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % AB_NST;
                mbar_wait(&bar_empty[s], ((kb / AB_NST) & 1) ^ 1);
                mbar_expect_tx(&bar_full[s], TMA_BYTES);
                uint8_t* st = smem + s * STAGE;
                tma_load_5d(st, &tmA, &bar_full[s], 0, i1 * Cf::KB, i2, i3, tile * Cf::G128);
                tma_load_5d(st + X_BYTES, &tmB1, &bar_full[s], 0, i1 * Cf::KB, i2, i3, 0);
                tma_load_5d(st + 2 * X_BYTES, &tmB2, &bar_full[s], 0, i1 * Cf::KB + g.sh1, i2 + g.sh2, i3, 0);
                if (++i1 == g.n1chunks) { i1 = 0; if (++i2 == g.D2) { i2 = 0; ++i3; } }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % AB_NST;
                mbar_wait(&bar_full[s], (kb / AB_NST) & 1);
                tc_fence_after();
                const uint32_t sx = smem_u32(smem + s * STAGE), sy = sx + X_BYTES, so = sx + TMA_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t acc = (kb | kk) ? 1u : 0u;
                    ad_umma<kElem>(tmem, ad_desc_mn<kElem>(sx + kk * Cf::KSTEP), ad_desc_mn<kElem>(sy + kk * Cf::KSTEP), 256, acc);
                    ad_umma<kElem>(tmem + 256, ad_desc_mn<kElem>(sx + kk * Cf::KSTEP), ad_desc_mn<kElem>(so + kk * Cf::KSTEP), 32, acc);
                }
                umma_commit(&bar_empty[s]);
            }
            umma_commit(&bar_done);
        }
        __syncwarp();
    } else {
        // epilogue warps 2..5 -> TMEM lane quadrant warp % 4; lane = column of the A tile, TMEM column = column of [B1 | B2 | 1]
        const int q = warp & 3;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        float* dst = partial + (((long)split * tiles + tile) * AD_ROWS) * 128 + q * 32 + lane;
        if (num_kb > 0) {
            mbar_wait(&bar_done, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int c0 = 0; c0 < 288; c0 += 32) {
            float v[32];
            if (num_kb > 0) {
                tmem_ld32(taddr + c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (c0 < 256) {
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[(long)(c0 + j) * 128] = v[j];
            } else {
                dst[(long)256 * 128] = v[0];            // the ones group: 32 identical columns, one is enough
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

// C1 / C2 / colsum (+)= sum over splits (fixed order: deterministic)
__global__ void atb_dual_reduce_kernel(const float* __restrict__ partial, int splits, int tiles, float* __restrict__ C1,
                                       long ldc1, float* __restrict__ C2, long ldc2, float* __restrict__ colsum,
                                       int accumulate, int accumulate_colsum) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long per_tile = (long)AD_ROWS * 128;
    if (e >= (long)tiles * per_tile) return;
    const int tile = (int)(e / per_tile), r = (int)(e % per_tile), y = r >> 7, x = r & 127;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(((long)k * tiles + tile) * AD_ROWS + y) * 128 + x];
    const long a = (long)tile * 128 + x;
    if (y < 256) {
        float* dst = y < 128 ? C1 + a * ldc1 + y : C2 + a * ldc2 + (y - 128);
        *dst = accumulate ? *dst + s : s;
    } else {
        colsum[a] = accumulate_colsum ? colsum[a] + s : s;
    }
}

static void atb_tc_plan(long M, int ycols, int* nyt, int* tiles, int* splits, long* rows_per_split) {
    *nyt = ycols % 256 == 0 ? 256 : ycols % 128 == 0 ? 128 : 64;       // 64: the N = 64 operand (input_size) as one tile
    *tiles = ycols / *nyt;
    long sp = 148 / *tiles;
    const long max_sp = (M + 4 * AB_KB - 1) / (4 * AB_KB);       // at least 4 stages of rows per split
    if (sp > max_sp) sp = max_sp;
    if (sp < 1) sp = 1;
    long rps = (M + sp - 1) / sp;
    rps = (rps + AB_KB - 1) / AB_KB * AB_KB;
    *rows_per_split = rps;
    *splits = (int)((M + rps - 1) / rps);
}

}  // namespace dprnn

using namespace dprnn;

extern "C" int dprnn_gemm_atb_tc_supported(int N1, int N2, long lda, long ldb) {
    const bool shape = (N1 == 128 && (N2 % 128 == 0 || N2 == 64) && N2 <= 4096) ||
                       (N2 == 128 && (N1 % 128 == 0 || N1 == 64) && N1 <= 4096);
    return shape && lda % 4 == 0 && ldb % 4 == 0 && N1 > 0 && N2 > 0;
}

extern "C" size_t dprnn_gemm_atb_tc_workspace_bytes(int N1, int N2) {
    return (size_t)148 * (size_t)(N1 > N2 ? N1 : N2) * 128 * sizeof(float);
}

extern "C" int dprnn_gemm_atb_tc(const float* A, long lda, const float* B, long ldb, float* C, long ldc, long M, int N1,
                                 int N2, int accumulate, void* workspace, void* stream) {
    DPRNN_CHECK_ARG(A && B && C && workspace && M > 0 && M < (1L << 31) && dprnn_gemm_atb_tc_supported(N1, N2, lda, ldb));
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)B) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    // X = the 128-column operand (MMA M role); prefer X = B so that C rows are written along x (coalesced)
    const bool x_is_b = N2 == 128;
    const float* X = x_is_b ? B : A;
    const float* Y = x_is_b ? A : B;
    const long ldx = x_is_b ? ldb : lda, ldy = x_is_b ? lda : ldb;
    const int ycols = x_is_b ? N1 : N2;
    int nyt, tiles, splits;
    long rps;
    atb_tc_plan(M, ycols, &nyt, &tiles, &splits, &rps);
    CUtensorMap tmX, tmY;
    // dims innermost first: 32 columns (one 128-byte line), rows, column groups
    const uint64_t dX[3] = {32, (uint64_t)M, 4}, sX[3] = {4, (uint64_t)ldx * 4, 128};
    const uint32_t bX[3] = {32, AB_KB, 4};
    const uint64_t dY[3] = {32, (uint64_t)M, (uint64_t)(ycols / 32)}, sY[3] = {4, (uint64_t)ldy * 4, 128};
    const uint32_t bY[3] = {32, AB_KB, (uint32_t)(nyt / 32)};
    if (make_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, X, dX, sX, bX, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    if (make_tmap(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, Y, dY, sY, bY, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    const size_t smem = (size_t)AB_NST * (4 + nyt / 32) * AB_GROUP + 1024;
    DPRNN_CUDA(cudaFuncSetAttribute(atb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    atb_tc_kernel<<<dim3(tiles, splits), 192, smem, st>>>(tmX, tmY, M, rps, nyt, ycols, (float*)workspace);
    DPRNN_CHECK_LAUNCH();
    const long total = (long)ycols * 128;
    atb_tc_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)workspace, splits, ycols, C, ldc,
                                                                        x_is_b ? 0 : 1, accumulate);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

// C[N1,128] (+)= A[M,N1]^T B[M,128] and colsum[N1] (+)= column sums of A, in one pass over A (see atb_tc_colsum_kernel).
extern "C" int dprnn_gemm_atb_tc_colsum_supported(int N1, int N2, long lda, long ldb) {
    return N2 == 128 && N1 % 128 == 0 && N1 > 0 && N1 <= 4096 && lda % 4 == 0 && ldb % 4 == 0;
}

extern "C" size_t dprnn_gemm_atb_tc_colsum_workspace_bytes(int N1) {
    return (size_t)(148 + N1 / 128) * (size_t)(32 * AB_YG) * 128 * sizeof(float);     // splits * tiles <= 148 (+ rounding)
}

extern "C" int dprnn_gemm_atb_tc_colsum(const float* A, long lda, const float* B, long ldb, float* C, long ldc, float* colsum,
                                        long M, int N1, int N2, int accumulate, int accumulate_colsum, void* workspace,
                                        void* stream) {
    DPRNN_CHECK_ARG(A && B && C && colsum && workspace && M > 0 && M < (1L << 31));
    DPRNN_CHECK_ARG(dprnn_gemm_atb_tc_colsum_supported(N1, N2, lda, ldb));
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)B) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int tiles = N1 / 128;
    long sp = 148 / tiles;
    const long max_sp = (M + 4 * AB_KB - 1) / (4 * AB_KB);
    if (sp > max_sp) sp = max_sp;
    if (sp < 1) sp = 1;
    long rps = (M + sp - 1) / sp;
    rps = (rps + AB_KB - 1) / AB_KB * AB_KB;
    const int splits = (int)((M + rps - 1) / rps);
    CUtensorMap tmX, tmY;
    const uint64_t dX[3] = {32, (uint64_t)M, (uint64_t)(N1 / 32)}, sX[3] = {4, (uint64_t)lda * 4, 128};
    const uint32_t bX[3] = {32, AB_KB, 4};
    const uint64_t dY[3] = {32, (uint64_t)M, 4}, sY[3] = {4, (uint64_t)ldb * 4, 128};
    const uint32_t bY[3] = {32, AB_KB, 4};
    if (make_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, A, dX, sX, bX, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    if (make_tmap(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, B, dY, sY, bY, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    const size_t smem = (size_t)AB_NST * (4 + AB_YG) * AB_GROUP + 1024;
    DPRNN_CUDA(cudaFuncSetAttribute(atb_tc_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    atb_tc_colsum_kernel<<<dim3(tiles, splits), 192, smem, st>>>(tmX, tmY, M, rps, tiles, (float*)workspace);
    DPRNN_CHECK_LAUNCH();
    const long total = (long)tiles * 129 * 128;
    atb_tc_colsum_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)workspace, splits, tiles, C, ldc,
                                                                               colsum, accumulate, accumulate_colsum);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

// One pass over A for C1 (+)= A^T B1, C2 (+)= A^T shift_t(B2), colsum (+)= column sums of A (see atb_dual_kernel).
// Rows are the chunk positions of a [B, S, K] batch: row = (b*S + s)*K + k.  inter = 0: sequences run along k (the intra-
// chunk layer, time = k); inter = 1: along s (time = s).  shift in {-1, 0, +1}: B2 is read at time t + shift, zero outside.
// is_bf16: all three operands bf16 (leading dimensions in elements, multiples of 64), else fp32 read as TF32.
extern "C" int dprnn_gemm_atb_dual_supported(int is_bf16, int N1, long lda, long ldb1, long ldb2) {
    const int gc = is_bf16 ? 64 : 32;
    return N1 > 0 && N1 % 128 == 0 && N1 <= 4096 && lda % gc == 0 && ldb1 % gc == 0 && ldb2 % gc == 0 && lda >= N1 &&
           ldb1 >= 128 && ldb2 >= 128;
}

extern "C" size_t dprnn_gemm_atb_dual_workspace_bytes(int N1) {
    return (size_t)(148 + N1 / 128) * (size_t)AD_ROWS * 128 * sizeof(float);      // splits * tiles <= 148 (+ rounding)
}

extern "C" int dprnn_gemm_atb_dual(const void* A, int is_bf16, long lda, int N1, const void* B1, long ldb1, const void* B2,
                                   long ldb2, int B, int S, int K, int inter, int shift, float* C1, long ldc1, float* C2,
                                   long ldc2, float* colsum, int accumulate, int accumulate_colsum, void* workspace,
                                   void* stream) {
    DPRNN_CHECK_ARG(A && B1 && B2 && C1 && C2 && colsum && workspace && B > 0 && S > 0 && K > 0);
    DPRNN_CHECK_ARG(dprnn_gemm_atb_dual_supported(is_bf16, N1, lda, ldb1, ldb2) && shift >= -1 && shift <= 1);
    DPRNN_CHECK_ARG(((uintptr_t)A | (uintptr_t)B1 | (uintptr_t)B2) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int tiles = N1 / 128;
    const int kb_rows = is_bf16 ? AdCfg<2>::KB : AdCfg<4>::KB, gcols = is_bf16 ? 64 : 32, g128 = is_bf16 ? 2 : 4;
    const uint64_t el = is_bf16 ? 2 : 4;
    // d1 = k (unit row stride) in both layouts; intra: d2 = (b, s) flattened, time = d1; inter: d2 = s = time, d3 = b
    const uint64_t D1 = (uint64_t)K, D2 = inter ? (uint64_t)S : (uint64_t)B * S, D3 = inter ? (uint64_t)B : 1;
    AtbDualGeo g;
    g.n1chunks = (int)((D1 + kb_rows - 1) / kb_rows);
    g.D2 = (int)D2;
    g.sh1 = inter ? 0 : shift;
    g.sh2 = inter ? shift : 0;
    g.total_chunks = (long)g.n1chunks * (long)D2 * (long)D3;
    long sp = 148 / tiles;
    const long max_sp = (g.total_chunks + 3) / 4;                 // at least 4 stages of rows per split
    if (sp > max_sp) sp = max_sp;
    if (sp < 1) sp = 1;
    g.chunks_per_split = (g.total_chunks + sp - 1) / sp;
    const int splits = (int)((g.total_chunks + g.chunks_per_split - 1) / g.chunks_per_split);
    auto make = [&](CUtensorMap* m, const void* base, long ld, int groups) {
        const uint64_t d[5] = {(uint64_t)gcols, D1, D2, D3, (uint64_t)groups};
        const uint64_t sb[5] = {el, (uint64_t)ld * el, (uint64_t)K * ld * el, (uint64_t)S * K * ld * el, 128};
        const uint32_t bx[5] = {(uint32_t)gcols, (uint32_t)kb_rows, 1, 1, (uint32_t)g128};
        return make_tmap(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, d, sb, bx,
                         is_bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    };
    CUtensorMap tmA, tmB1, tmB2;
    if (make(&tmA, A, lda, N1 / gcols) || make(&tmB1, B1, ldb1, g128) || make(&tmB2, B2, ldb2, g128)) return 1;
    const size_t smem = is_bf16 ? AdCfg<2>::SMEM : AdCfg<4>::SMEM;
    auto kern = is_bf16 ? atb_dual_kernel<2> : atb_dual_kernel<4>;
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(tiles, splits), 192, smem, st>>>(tmA, tmB1, tmB2, g, tiles, (float*)workspace);
    DPRNN_CHECK_LAUNCH();
    const long total = (long)tiles * AD_ROWS * 128;
    atb_dual_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)workspace, splits, tiles, C1, ldc1,
                                                                          C2, ldc2, colsum, accumulate, accumulate_colsum);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
