// Linear -> GroupNorm(1,F) / gLN -> residual add as ONE persistent tensor-core kernel (bf16 mode), i.e. the tail of
// every DPRNN half-block (dprnn.py:86-92, 96-99):
//     y = h W^T + b            h [M,K] bf16 (LSTM output), W [128,K] bf16
//     x += (y - mean_u) * rstd_u * gamma + beta         per utterance u; x fp32 [M,128] in place, bf16 shadow xb
// The norm needs the statistics of the WHOLE utterance before any output can be written, so the op is two passes over
// h.  Neither pass writes y: pass 0 recomputes nothing but reduces {sum, sumsq} of the accumulators, pass 1 recomputes
// the (cheap, 64-128 KFLOP/row) product and applies norm + residual in its epilogue.  Both passes run inside one launch:
// work items (pass, 128-row tile) are handed out through an atomic ticket in an order where pass 0 leads pass 1 by
// `lead` tiles (>= the longest utterance), so
//   * the second read of an h tile comes ~lead tiles after the first and is served by the 126 MB L2 (h loads of pass 0
//     carry an evict_last hint, every other stream evict_first): HBM sees  h + x(in) + x(out) + xb  = 3.5 A per
//     half-block instead of  h + y + y + x + x + xb = 5.5 A  for Linear-then-norm kernels;
//   * a pass-1 item only ever waits (spin on a per-utterance flag) for items with SMALLER tickets, which are owned by
//     resident CTAs - no co-residency assumption, no deadlock with other kernels on other streams.
// Statistics are deterministic: every pass-0 tile stores its partial sums; the CTA that completes an utterance
// reduces the utterance's partials in a fixed order.
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

constexpr int LN_N = 128;
constexpr uint32_t LN_BLK = 128 * 128;      // one [128 rows x 128 B] swizzled tile = 16 KiB
constexpr int LN_AST = 4;                   // h K-block ring
constexpr int LN_XST = 4;                   // x chunk ring ([128 rows x 32 fp32])
constexpr int LN_IST = 8;                   // work-item ring
constexpr int LN_THREADS = 416;

struct LnItem { int pass, tile, u0, bnd; float2 mr0, mr1; };

struct LnArgs {
    const float *bias, *gamma, *beta;
    const long* row_off;       // [n_utt + 1] first row of every utterance (row_off[n_utt] = M)
    int n_utt, M, tiles, lead;
    double eps;
    unsigned* ticket;          // workspace, zeroed before launch
    unsigned* utt_count;       // [n_utt] pass-0 tiles that have contributed
    unsigned* utt_flag;        // [n_utt] 1 once mean_rstd[u] is published
    double2* partial;          // [tiles][2] {sum, sumsq} of the rows before / after the utterance boundary in the tile
    float2* mean_rstd;         // [n_utt]
};

// shared-memory control block, placed after the tiles
struct LnCtl {
    uint64_t a_full[LN_AST], a_empty[LN_AST], x_full[LN_XST], x_empty[LN_XST], w_full, acc_full[2], acc_empty[2],
        i_full[LN_IST], i_empty[LN_IST], red_full[2], red_empty[2];
    LnItem items[LN_IST];
    double red[2][8][4];
    int red_tile[2][2];        // {tile, u0 | bnd << 20}... kept simple: [buf] = {tile, item slot}
    float bias[LN_N], gamma[LN_N], beta[LN_N];
    uint32_t tmem_base;
};
constexpr uint32_t LN_SMEM_TILES = (4 + LN_AST + LN_XST + 2) * LN_BLK;     // KB=4 worst case

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void ln_tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void ln_tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(m), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ticket -> (pass, tile): pass 0 runs `lead` tiles ahead, then the passes alternate, then pass 1 drains
__device__ __forceinline__ void ln_decode(unsigned t, int tiles, int lead, int& pass, int& tile) {
    if (lead >= tiles) { pass = t >= (unsigned)tiles; tile = pass ? (int)t - tiles : (int)t; return; }
    if (t < (unsigned)lead) { pass = 0; tile = (int)t; return; }
    const int r = (int)t - lead, n_inter = tiles - lead;
    if (r < 2 * n_inter) { pass = r & 1; tile = pass ? (r >> 1) : lead + (r >> 1); return; }
    pass = 1; tile = r - n_inter;
}

// Warp roles: 0 = ticket scheduler (decodes items, fetches the statistics a pass-1 item needs), 1 = MMA issuer,
// 2 = statistics bookkeeping (global atomics, the per-utterance reduction), 3 = TMA producer for h, 12 = TMA producer
// for x (separate threads, so neither stream of loads queues behind the other's ring), 4..11 = epilogue: TMEM lane
// quadrant = warp % 4, column half = (warp - 4) / 4.
template <int KB>     // K / 64
__global__ void __launch_bounds__(LN_THREADS, 1) linear_norm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmW,
                                                                    const __grid_constant__ CUtensorMap tmX,
                                                                    const __grid_constant__ CUtensorMap tmXb,
                                                                    const LnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sW = smem;                                 // KB blocks
    uint8_t* sA = sW + KB * LN_BLK;                     // LN_AST blocks
    uint8_t* sX = sA + LN_AST * LN_BLK;                 // LN_XST chunk buffers
    uint8_t* sB = sX + LN_XST * LN_BLK;                 // one bf16 staging tile [128 rows x 64 bf16] per column half
    LnCtl& ctl = *reinterpret_cast<LnCtl*>(sB + 2 * LN_BLK);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) { printf("linear_norm_kernel: shared memory base not 1024-byte aligned\n"); __trap(); }
        prefetch_tmap(&tmA); prefetch_tmap(&tmW); prefetch_tmap(&tmX); prefetch_tmap(&tmXb);
        for (int s = 0; s < LN_AST; ++s) { mbar_init(&ctl.a_full[s], 1); mbar_init(&ctl.a_empty[s], 1); }
        for (int s = 0; s < LN_XST; ++s) { mbar_init(&ctl.x_full[s], 1); mbar_init(&ctl.x_empty[s], 1); }
        for (int s = 0; s < LN_IST; ++s) { mbar_init(&ctl.i_full[s], 1); mbar_init(&ctl.i_empty[s], 12); }
        mbar_init(&ctl.w_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&ctl.acc_full[s], 1); mbar_init(&ctl.acc_empty[s], 8);
            mbar_init(&ctl.red_full[s], 8); mbar_init(&ctl.red_empty[s], 1);
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < LN_N; i += blockDim.x) {
        ctl.bias[i] = a.bias[i]; ctl.gamma[i] = a.gamma[i]; ctl.beta[i] = a.beta[i];
    }
    if (warp == 1) tmem_alloc<1>(&ctl.tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = ctl.tmem_base;
    const unsigned total = 2u * (unsigned)a.tiles;

    if (warp == 0) {
        // ================= scheduler + TMA producer =================
        if (elect_one()) {
            mbar_expect_tx(&ctl.w_full, KB * LN_BLK);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * LN_BLK, &tmW, &ctl.w_full, kb * 64, 0);
            unsigned tbase = 0;
            for (int n = 0;; ++n) {
                if ((n & 1) == 0) tbase = atomicAdd(a.ticket, 2u);      // two consecutive tickets per atomic
                const unsigned t = tbase + (n & 1);
                const int is = n % LN_IST;
                mbar_wait(&ctl.i_empty[is], ((n / LN_IST) & 1) ^ 1);
                LnItem it;
                it.mr0 = it.mr1 = make_float2(0.f, 0.f);
                if (t >= total) {
                    it.pass = -1; it.tile = it.u0 = it.bnd = 0;
                    ctl.items[is] = it;
                    mbar_arrive(&ctl.i_full[is]);
                    break;
                }
                ln_decode(t, a.tiles, a.lead, it.pass, it.tile);
                {   // utterance of the tile's first row, and how many of its 128 rows belong to it
                    const long row0 = (long)it.tile * 128;
                    int lo = 0, hi = a.n_utt - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (__ldg(a.row_off + mid) <= row0) lo = mid; else hi = mid - 1;
                    }
                    it.u0 = lo;
                    const long rest = __ldg(a.row_off + lo + 1) - row0;
                    it.bnd = rest < 128 ? (int)rest : 128;
                }
                if (it.pass == 1) {
                    // statistics of the utterance(s) under this tile: published by items with smaller tickets
                    const bool need1 = it.bnd < 128 && it.u0 + 1 < a.n_utt;
                    long spins = 0;
                    while (!ld_acquire_u32(a.utt_flag + it.u0) || (need1 && !ld_acquire_u32(a.utt_flag + it.u0 + 1))) {
                        __nanosleep(64);
                        if (++spins > (1L << 25)) { printf("linear_norm_kernel: statistics of utterance %d never arrived\n", it.u0); __trap(); }
                    }
                    it.mr0 = __ldcg(a.mean_rstd + it.u0);
                    if (need1) it.mr1 = __ldcg(a.mean_rstd + it.u0 + 1);
                }
                ctl.items[is] = it;
                mbar_arrive(&ctl.i_full[is]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, LN_N);
            mbar_wait(&ctl.w_full, 0);
            int ait = 0;
            for (int n = 0;; ++n) {
                const int is = n % LN_IST;
                mbar_wait(&ctl.i_full[is], (n / LN_IST) & 1);
                const int pass = ctl.items[is].pass;
                mbar_arrive(&ctl.i_empty[is]);
                if (pass < 0) break;
                const int acc = n & 1;
                mbar_wait(&ctl.acc_empty[acc], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb, ++ait) {
                    const int s = ait % LN_AST;
                    mbar_wait(&ctl.a_full[s], (ait / LN_AST) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(sA + s * LN_BLK), sb = smem_u32(sW + kb * LN_BLK);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16<1>(tmem + acc * LN_N, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sb + kk * 32), idesc,
                                     (kb | kk) ? 1u : 0u);
                    umma_commit(&ctl.a_empty[s]);
                }
                umma_commit(&ctl.acc_full[acc]);
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ================= statistics bookkeeping =================
        // For every pass-0 item: publish the tile's partial sums, count the tile towards its utterance(s) and, when it
        // completed one, reduce that utterance's partials in tile order (deterministic) and raise its flag.  Global
        // round trips (fence, atomic) stay off the epilogue warps' critical path.
        int n0 = 0;
        for (int n = 0;; ++n) {
            const int is = n % LN_IST;
            mbar_wait(&ctl.i_full[is], (n / LN_IST) & 1);
            const LnItem it = ctl.items[is];
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl.i_empty[is]);
            if (it.pass < 0) break;
            if (it.pass != 0) continue;
            const int rb = n0 & 1;
            mbar_wait(&ctl.red_full[rb], (n0 >> 1) & 1);
            int fin0 = 0, fin1 = 0;
            if (lane == 0) {
                double2 p0 = make_double2(0.0, 0.0), p1 = make_double2(0.0, 0.0);
                for (int w = 0; w < 8; ++w) {       // fixed order
                    p0.x += ctl.red[rb][w][0]; p0.y += ctl.red[rb][w][1];
                    p1.x += ctl.red[rb][w][2]; p1.y += ctl.red[rb][w][3];
                }
                mbar_arrive(&ctl.red_empty[rb]);
                a.partial[2 * (long)it.tile] = p0;
                a.partial[2 * (long)it.tile + 1] = p1;
                __threadfence();
                for (int k = 0; k < 2; ++k) {
                    const int u = it.u0 + k;
                    if (k == 1 && (it.bnd >= 128 || u >= a.n_utt)) break;
                    const long f = __ldg(a.row_off + u), l = __ldg(a.row_off + u + 1) - 1;
                    const unsigned expected = (unsigned)(l / 128 - f / 128 + 1);
                    const unsigned old = atomicAdd(a.utt_count + u, 1u);
                    if (old + 1 == expected) { if (k == 0) fin0 = 1; else fin1 = 1; }
                }
                __threadfence();
            }
            ++n0;
            fin0 = __shfl_sync(0xffffffffu, fin0, 0);
            fin1 = __shfl_sync(0xffffffffu, fin1, 0);
            for (int k = 0; k < 2; ++k) {
                if (!(k == 0 ? fin0 : fin1)) continue;
                const int u = it.u0 + k;
                const long f = __ldg(a.row_off + u), e = __ldg(a.row_off + u + 1);
                const long t0 = f / 128, t1 = (e - 1) / 128;
                double s = 0.0, qq = 0.0;
                for (long tt = t0 + lane; tt <= t1; tt += 32) {
                    // the utterance's first tile may start inside the previous utterance: then its rows sit in slot 1
                    const int slot = (tt * 128 < f) ? 1 : 0;
                    const double2 p = __ldcg(a.partial + 2 * tt + slot);
                    s += p.x; qq += p.y;
                }
                s = warp_sum(s); qq = warp_sum(qq);
                if (lane == 0) {
                    const double cnt = (double)(e - f) * LN_N;
                    const double mean = s / cnt;
                    double var = qq / cnt - mean * mean;
                    if (var < 0.0) var = 0.0;
                    a.mean_rstd[u] = make_float2((float)mean, (float)(1.0 / sqrt(var + a.eps)));
                    __threadfence();
                    st_release_u32(a.utt_flag + u, 1u);
                }
                __syncwarp();
            }
        }
    } else if (warp == 3) {
        // ================= h producer: K-blocks of every item into the A ring =================
        if (elect_one()) {
            const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
            int ait = 0;
            for (int n = 0;; ++n) {
                const int is = n % LN_IST;
                mbar_wait(&ctl.i_full[is], (n / LN_IST) & 1);
                const int pass = ctl.items[is].pass, tile = ctl.items[is].tile;
                mbar_arrive(&ctl.i_empty[is]);
                if (pass < 0) break;
                const uint64_t pol = pass == 0 ? pol_keep : pol_stream;      // the second read of h is the last one
                for (int kb = 0; kb < KB; ++kb, ++ait) {
                    const int s = ait % LN_AST;
                    mbar_wait(&ctl.a_empty[s], ((ait / LN_AST) & 1) ^ 1);
                    mbar_expect_tx(&ctl.a_full[s], LN_BLK);
                    ln_tma_load_2d(sA + s * LN_BLK, &tmA, &ctl.a_full[s], kb * 64, tile * 128, pol);
                }
            }
        }
        __syncwarp();
    } else if (warp == 12) {
        // ================= x producer: the four 32-column chunks of every pass-1 item =================
        if (elect_one()) {
            const uint64_t pol_stream = l2_policy_evict_first();
            int xit = 0;
            for (int n = 0;; ++n) {
                const int is = n % LN_IST;
                mbar_wait(&ctl.i_full[is], (n / LN_IST) & 1);
                const int pass = ctl.items[is].pass, tile = ctl.items[is].tile;
                mbar_arrive(&ctl.i_empty[is]);
                if (pass < 0) break;
                if (pass != 1) continue;
                for (int c = 0; c < 4; ++c, ++xit) {
                    const int s = xit % LN_XST;
                    mbar_wait(&ctl.x_empty[s], ((xit / LN_XST) & 1) ^ 1);
                    mbar_expect_tx(&ctl.x_full[s], LN_BLK);
                    ln_tma_load_2d(sX + s * LN_BLK, &tmX, &ctl.x_full[s], c * 32, tile * 128, pol_stream);
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int q = warp & 3, ew = warp - 4, half = ew >> 2;
        const int r = q * 32 + lane;                 // row inside the tile = TMEM lane
        const bool storer = (q == 0 && lane == 0);   // one per column half
        const uint64_t pol_stream = l2_policy_evict_first();
        uint8_t* bbuf = sB + half * LN_BLK;
        int n0 = 0, n1 = 0;
        for (int n = 0;; ++n) {
            const int is = n % LN_IST;
            mbar_wait(&ctl.i_full[is], (n / LN_IST) & 1);
            const LnItem it = ctl.items[is];
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl.i_empty[is]);
            if (it.pass < 0) break;
            const int acc = n & 1;
            const long row = (long)it.tile * 128 + r;
            const bool valid = row < a.M;
            const bool second = r >= it.bnd;          // row belongs to utterance u0 + 1
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * LN_N + half * 64;
            mbar_wait(&ctl.acc_full[acc], (n >> 1) & 1);
            tc_fence_after();

            if (it.pass == 0) {
                float s_sum = 0.f, s_sq = 0.f;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    float v[32];
                    tmem_ld32(taddr + cc * 32, v);
                    if (cc == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&ctl.acc_empty[acc]);
                    }
                    float s = 0.f, qq = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float y = v[j] + ctl.bias[half * 64 + cc * 32 + j];
                        s += y; qq = fmaf(y, y, qq);
                    }
                    s_sum += s; s_sq += qq;
                }
                const double ds = valid ? (double)s_sum : 0.0, dq = valid ? (double)s_sq : 0.0;
                const double w0 = warp_sum(second ? 0.0 : ds), w1 = warp_sum(second ? 0.0 : dq);
                const double w2 = warp_sum(second ? ds : 0.0), w3 = warp_sum(second ? dq : 0.0);
                const int rb = n0 & 1;
                mbar_wait(&ctl.red_empty[rb], ((n0 >> 1) & 1) ^ 1);
                if (lane == 0) {
                    ctl.red[rb][ew][0] = w0; ctl.red[rb][ew][1] = w1; ctl.red[rb][ew][2] = w2; ctl.red[rb][ew][3] = w3;
                    mbar_arrive(&ctl.red_full[rb]);
                }
                ++n0;
                continue;
            }

            // ---------------- pass 1: norm + residual on this warp group's 64 columns ----------------
            const float mean = second ? it.mr1.x : it.mr0.x, rstd = second ? it.mr1.y : it.mr0.y;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c = half * 2 + cc, c0 = c * 32;
                float v[32];
                tmem_ld32(taddr + cc * 32, v);
                if (cc == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ctl.acc_empty[acc]);
                }
                const int xi = n1 * 4 + c, xs = xi % LN_XST;
                uint8_t* xbuf = sX + xs * LN_BLK;
                mbar_wait(&ctl.x_full[xs], (xi / LN_XST) & 1);
                if (cc == 0) asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");     // staging tile free (the storer drained its stores at the end of the previous item)
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4* px = reinterpret_cast<float4*>(xbuf + sw128_offset(r, j));
                    float4 xv = *px;
                    const int k0 = c0 + 4 * j;
                    xv.x += fmaf((v[4 * j + 0] + ctl.bias[k0 + 0] - mean) * rstd, ctl.gamma[k0 + 0], ctl.beta[k0 + 0]);
                    xv.y += fmaf((v[4 * j + 1] + ctl.bias[k0 + 1] - mean) * rstd, ctl.gamma[k0 + 1], ctl.beta[k0 + 1]);
                    xv.z += fmaf((v[4 * j + 2] + ctl.bias[k0 + 2] - mean) * rstd, ctl.gamma[k0 + 2], ctl.beta[k0 + 2]);
                    xv.w += fmaf((v[4 * j + 3] + ctl.bias[k0 + 3] - mean) * rstd, ctl.gamma[k0 + 3], ctl.beta[k0 + 3]);
                    *px = xv;
                    __nv_bfloat162 lo = __floats2bfloat162_rn(xv.x, xv.y), hi = __floats2bfloat162_rn(xv.z, xv.w);
                    pk[2 * j] = *reinterpret_cast<uint32_t*>(&lo);
                    pk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&hi);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(bbuf + sw128_offset(r, cc * 4 + j)) =
                        make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                fence_async_smem();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");          // chunk complete in shared memory
                if (storer) {
                    ln_tma_store_2d(&tmX, xbuf, c0, it.tile * 128, pol_stream);
                    if (cc == 1) ln_tma_store_2d(&tmXb, bbuf, half * 64, it.tile * 128, pol_stream);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (cc == 1) {
                        // both chunks and the bf16 tile have been read out of shared memory: hand the two x buffers
                        // back to the producer right away so that the next item's x loads start early
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        mbar_arrive(&ctl.x_empty[(xi - 1) % LN_XST]);
                        mbar_arrive(&ctl.x_empty[xi % LN_XST]);
                    }
                }
            }
            ++n1;
        }
        if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, 256);
}

template <int KB>
static int launch_ln(const void* A, const void* W, float* X, void* Xb, const LnArgs& args, cudaStream_t st) {
    constexpr int K = KB * 64;
    const int M = args.M;
    CUtensorMap tmA, tmW, tmX, tmXb;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {2, (uint64_t)K * 2};
    const uint32_t bA[2] = {64, 128};
    const uint64_t dW[2] = {(uint64_t)K, (uint64_t)LN_N}, sW[2] = {2, (uint64_t)K * 2};
    const uint32_t bW[2] = {64, (uint32_t)LN_N};
    const uint64_t dX[2] = {(uint64_t)LN_N, (uint64_t)M}, sX[2] = {4, (uint64_t)LN_N * 4};
    const uint32_t bX[2] = {32, 128};
    const uint64_t dB[2] = {(uint64_t)LN_N, (uint64_t)M}, sB[2] = {2, (uint64_t)LN_N * 2};
    const uint32_t bB[2] = {64, 128};
    if (make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dA, sA, bA)) return 1;
    if (make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dW, sW, bW)) return 1;
    if (make_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, X, dX, sX, bX)) return 1;
    if (make_tmap(&tmXb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Xb, dB, sB, bB)) return 1;
    const size_t smem = (size_t)(KB + LN_AST + LN_XST + 2) * LN_BLK + sizeof(LnCtl);
    static_assert((4 + LN_AST + LN_XST + 2) * LN_BLK + sizeof(LnCtl) <= 232448, "shared memory budget of one SM (227 KiB)");
    auto kern = linear_norm_kernel<KB>;
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = 2 * args.tiles;
    kern<<<items < sms ? items : sms, LN_THREADS, smem, st>>>(tmA, tmW, tmX, tmXb, args);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

static size_t ln_align(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace dprnn

using namespace dprnn;

extern "C" size_t dprnn_linear_norm_workspace_bytes(int M, int n_utt) {
    const size_t tiles = (size_t)cdiv(M, 128);
    return 256 + 2 * ln_align((size_t)n_utt * 4) + ln_align(tiles * 2 * sizeof(double2)) + ln_align((size_t)n_utt * sizeof(float2));
}

extern "C" int dprnn_linear_norm_residual_bf16(const void* h, const void* W, const float* bias, float* x, void* x_bf16,
                                               const float* gamma, const float* beta, float eps, const long* row_off,
                                               int n_utt, long max_rows_per_utt, int M, int K, void* workspace,
                                               void* stream) {
    DPRNN_CHECK_ARG(h && W && bias && x && x_bf16 && gamma && beta && row_off && workspace);
    DPRNN_CHECK_ARG(M > 0 && n_utt > 0 && max_rows_per_utt > 0 && (K == 128 || K == 256));
    DPRNN_CHECK_ARG(((uintptr_t)h | (uintptr_t)W | (uintptr_t)x | (uintptr_t)x_bf16 | (uintptr_t)workspace) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    LnArgs a;
    a.bias = bias; a.gamma = gamma; a.beta = beta; a.row_off = row_off;
    a.n_utt = n_utt; a.M = M; a.tiles = (int)cdiv(M, 128);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // pass 0 must lead by at least one utterance (+ one tile for a misaligned start); one more wave of slack keeps
    // pass-1 items from ever catching up with the statistics they need
    a.lead = (int)cdiv(max_rows_per_utt, 128) + 2 + sms;
    a.eps = (double)eps;
    const size_t zero_bytes = 256 + 2 * ln_align((size_t)n_utt * 4);
    a.ticket = (unsigned*)ws;
    a.utt_count = (unsigned*)(ws + 256);
    a.utt_flag = (unsigned*)(ws + 256 + ln_align((size_t)n_utt * 4));
    a.partial = (double2*)(ws + zero_bytes);
    a.mean_rstd = (float2*)(ws + zero_bytes + ln_align((size_t)a.tiles * 2 * sizeof(double2)));
    DPRNN_CUDA(cudaMemsetAsync(ws, 0, zero_bytes, st));
    return K == 256 ? launch_ln<4>(h, W, x, x_bf16, a, st) : launch_ln<2>(h, W, x, x_bf16, a, st);
}
