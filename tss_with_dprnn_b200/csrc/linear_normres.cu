// The tail of a DPRNN half-block (dprnn.py:86-92 / 96-99) as ONE persistent kernel that moves the Linear output through L2
// only:   y = h W^T + b  (tcgen05, 16-bit out, fp32 statistics)   then   xb <- xb + norm_u(y)   (GroupNorm(1,128) / gLN)
//
// Two kernels (linear_persist_kernel, then norm_residual_bf16res_kernel) move 4.76 GB per half-block at B = 64: read hb
// 1.59, write y 0.79 | read y 0.79, read xb 0.79, write xb 0.79.  The norm needs the statistics of the WHOLE utterance, so
// y has to exist somewhere between the two passes - but only one utterance's worth (12.4 MB of 16-bit y) at a time, which
// the 126 MB L2 holds.  Here both passes run inside one launch:
//   * warps 0..5 are linear_persist_kernel unchanged (TMA ring for hb with an evict_first hint, W resident, two TMEM
//     accumulators, swizzled staging + TMA stores of y, per-row {sum, sumsq}); tiles are drawn by ticket in row order, i.e.
//     utterance by utterance; the storer publishes a tile's rows in a per-utterance counter once its stores have landed;
//   * warps 6..13 are 8 independent pass-1 workers drawing items by a second ticket, also in utterance order: item (u, 0)
//     waits for the counter of u to reach its row count, reduces the per-row sums with the arithmetic of
//     row_stats_finalize_kernel (one warp playing its 256 threads: results are bit-identical to the two-kernel path and to
//     the ragged path), publishes mean / rstd and a flag; items (u, c > 0) wait for that flag and apply norm + residual to
//     128 rows with the arithmetic of norm_residual_bf16res_kernel (norm_res1).  y is read ~one utterance after it was written - an L2 hit - and, with
//     `discard`, dropped from L2 afterwards (discard.global.L2) so that it is never written back to HBM either.
// A pass-1 item only ever waits for work with smaller tickets, owned by resident CTAs: no deadlock, whatever else runs.
// HBM traffic: read hb 1.59 + read xb 0.79 + write xb 0.79 = 3.18 GB (+ 0.79 if y's dirty lines are written back).
#include "tc_common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {
using namespace tc;

constexpr int LR_N = 128, LR_AST = 6, LR_CST = 2, LR_TQ = 4;
constexpr uint32_t LR_BLK = 128 * 128;                 // one [128 rows x 128 B] swizzled tile = 16 KiB
constexpr int LR_CH = 128;                             // rows per pass-1 item
constexpr int LR_THREADS = 192 + 256;

struct LnrArgs {
    const float *bias, *gamma, *beta;
    float2* stats;             // [M] per-row {sum, sumsq}
    int M, tiles, n_utt, discard, lead;    // lead > 0: pass 0 never runs more than `lead` utterances ahead of pass 1
    long rows_per_utt;
    double eps;
    unsigned* ticket;          // pass-0 tile ticket           } workspace, zeroed before the launch
    unsigned* ticket1;         // pass-1 item ticket            }
    unsigned* rows_done;       // [n_utt] rows of the utterance whose y / sums are in global memory
    unsigned* flag;            // [n_utt] 1 once mean_rstd[u] is published
    unsigned* parts_done;      // [n_utt] statistics parts finished
    double* part8;             // [n_utt][16] the 8 warp results of {sum, sumsq}
    unsigned* chunks_done;     // [n_utt] pass-1 chunks finished (flow control only)
    float* mean_rstd;          // [n_utt, 2]
    const uint4* y;            // [M, 128] 16-bit (the buffer the TMA stores of pass 0 fill)
    uint4* xb;                 // [M, 128] 16-bit residual stream, updated in place
};

__device__ __forceinline__ void lr_tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(m), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t lr_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t lr_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void lr_tma_load_2d_hint(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned lr_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lr_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int KB, bool kF16>     // K / 64; 16-bit format of operands, y and the residual stream
__global__ void __launch_bounds__(LR_THREADS, 1) linear_normres_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                      const __grid_constant__ CUtensorMap tmW,
                                                                      const __grid_constant__ CUtensorMap tmC,
                                                                      const LnrArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;                                 // KB blocks
    uint8_t* sA = sW + KB * LR_BLK;                     // LR_AST blocks
    uint8_t* sC = sA + LR_AST * LR_BLK;                 // LR_CST staging blocks [128 rows x 64 16-bit]
    __shared__ __align__(8) uint64_t a_full[LR_AST], a_empty[LR_AST], w_full, acc_full[2], acc_empty[2], tq_full[LR_TQ],
        tq_empty[LR_TQ];
    __shared__ int tile_q[LR_TQ];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmW); prefetch_tmap(&tmC);
        for (int s = 0; s < LR_AST; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        mbar_init(&w_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        for (int s = 0; s < LR_TQ; ++s) { mbar_init(&tq_full[s], 1); mbar_init(&tq_empty[s], 5); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(&tmem_base_s, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const long R = a.rows_per_utt;
    const unsigned nch_all = (unsigned)((R + LR_CH - 1) / LR_CH);      // pass-1 chunks per utterance

    if (warp == 0) {
        if (elect_one()) {
            const uint64_t pol = lr_policy_evict_first();     // hb is read once
            mbar_expect_tx(&w_full, KB * LR_BLK);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * LR_BLK, &tmW, &w_full, kb * 64, 0);
            int it = 0;
            for (int n = 0;; ++n) {
                int tile = (int)atomicAdd(a.ticket, 1u);
                if (tile >= a.tiles) tile = -1;
                if (tile >= 0 && a.lead > 0) {
                    // flow control: y of at most lead + 1 utterances is alive at any time, so that it stays in L2 until pass 1
                    // has used it.  Before waiting, a flush entry makes the epilogue publish the tiles it still holds back
                    // (their rows may be what pass 1 of the awaited utterance is waiting for).
                    const long u = ((long)tile * 128) / R - a.lead;
                    if (u >= 0 && lr_ld_acquire(a.chunks_done + u) < nch_all) {
                        const int qf = n % LR_TQ;
                        mbar_wait(&tq_empty[qf], ((n / LR_TQ) & 1) ^ 1);
                        tile_q[qf] = -2;
                        mbar_arrive(&tq_full[qf]);
                        ++n;
                        while (lr_ld_acquire(a.chunks_done + u) < nch_all) __nanosleep(200);
                    }
                }
                const int qs = n % LR_TQ;
                mbar_wait(&tq_empty[qs], ((n / LR_TQ) & 1) ^ 1);
                tile_q[qs] = tile;
                mbar_arrive(&tq_full[qs]);
                if (tile < 0) break;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % LR_AST;
                    mbar_wait(&a_empty[s], ((it / LR_AST) & 1) ^ 1);
                    mbar_expect_tx(&a_full[s], LR_BLK);
                    lr_tma_load_2d_hint(sA + s * LR_BLK, &tmA, &a_full[s], kb * 64, tile * 128, pol);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_h16(128, LR_N, kF16);
            mbar_wait(&w_full, 0);
            int it = 0, m = 0;                                       // m counts real tiles (the accumulators alternate on it)
            for (int n = 0;; ++n) {
                const int qs = n % LR_TQ;
                mbar_wait(&tq_full[qs], (n / LR_TQ) & 1);
                const int tile = tile_q[qs];
                mbar_arrive(&tq_empty[qs]);
                if (tile == -2) continue;                            // flush entry: nothing to multiply
                if (tile < 0) break;
                const int acc = m & 1;
                mbar_wait(&acc_empty[acc], ((m >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                ++m;
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % LR_AST;
                    mbar_wait(&a_full[s], (it / LR_AST) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(sA + s * LR_BLK), sb = smem_u32(sW + kb * LR_BLK);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16<1>(tmem + acc * LR_N, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sb + kk * 32), idesc,
                                     (kb | kk) ? 1u : 0u);
                    umma_commit(&a_empty[s]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        // ================= pass 0 epilogue: y tile (16-bit) + per-row sums; the storer publishes finished tiles =================
        const int q = warp & 3;
        const int r_in_tile = q * 32 + lane;
        const bool storer = (warp == 2 && lane == 0);
        const uint64_t pol_y = lr_policy_evict_last();     // y is read again one utterance later: keep it in L2 until then
        constexpr int LAG = 3;                             // a tile is published once LAG younger tiles have been committed
        int chunk_it = 0, m = 0, hist[LAG], nhist = 0;     // m: real tiles; the storer's ring of committed, unpublished tiles
#pragma unroll
        for (int i = 0; i < LAG; ++i) hist[i] = -1;
        auto publish = [&](int tile) {                  // rows of `tile` -> the counter(s) of the utterance(s) it covers
            const long r0 = (long)tile * 128, r1 = r0 + 128 < a.M ? r0 + 128 : (long)a.M;
            const long u0 = r0 / R, split = (u0 + 1) * R < r1 ? (u0 + 1) * R : r1;
            asm volatile("fence.proxy.async;" ::: "memory");       // the TMA stores (async proxy) before generic-proxy readers
            __threadfence();
            atomicAdd(a.rows_done + u0, (unsigned)(split - r0));
            if (split < r1) atomicAdd(a.rows_done + u0 + 1, (unsigned)(r1 - split));
        };
        for (int n = 0;; ++n) {
            const int qs = n % LR_TQ;
            mbar_wait(&tq_full[qs], (n / LR_TQ) & 1);
            const int tile = tile_q[qs];
            __syncwarp();
            if (lane == 0) mbar_arrive(&tq_empty[qs]);
            if (tile == -2) {                               // flush entry: publish every tile still held back
                asm volatile("bar.sync 1, 128;" ::: "memory");     // every epilogue thread's sums of those tiles are written
                if (storer) {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    for (int i = 0; i < nhist; ++i) publish(hist[i]);
                    nhist = 0;
                }
                continue;
            }
            if (tile < 0) break;
            const int acc = m & 1;
            const long row = (long)tile * 128 + r_in_tile;
            mbar_wait(&acc_full[acc], (m >> 1) & 1);
            ++m;
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * LR_N;
            float s_sum = 0.f, s_sq = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < LR_N; c0 += 64, ++chunk_it) {
                uint32_t pk[32];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float v[32];
                    tmem_ld32(taddr + c0 + hh * 32, v);
                    if (c0 + 64 == LR_N && hh == 1) {       // accumulator fully read: hand it back to the MMA warp
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
                uint8_t* stage = sC + (chunk_it % LR_CST) * LR_BLK;
                if (storer) {
                    // the staging buffer has been READ by the TMA engine (cheap); a tile is published only when its stores
                    // have LANDED, which is waited for LAG tiles later so that it never stalls the store pipeline
                    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(LR_CST - 1) : "memory");
                    if (c0 == 0 && nhist == LAG) {
                        asm volatile("cp.async.bulk.wait_group %0;" ::"n"(2 * (LAG - 1)) : "memory");   // all but the newest LAG-1 tiles
                        publish(hist[0]);
#pragma unroll
                        for (int i = 0; i + 1 < LAG; ++i) hist[i] = hist[i + 1];
                        --nhist;
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(stage + sw128_offset(r_in_tile, j)) =
                        make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                fence_async_smem();
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (storer) {
                    lr_tma_store_2d(&tmC, stage, c0, tile * 128, pol_y);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            // (the storer passes a CTA barrier after these stores and fences at GPU scope before it publishes the tile)
            if (row < a.M) a.stats[row] = make_float2(s_sum, s_sq);
            if (storer) hist[nhist++] = tile;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // every epilogue thread's last sums are fenced
        if (storer) {
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            for (int i = 0; i < nhist; ++i) publish(hist[i]);
        }
    } else {
        // ================= pass 1: statistics of an utterance, then norm + residual on its rows =================
        // Every one of the 8 warps is an independent worker with its own tickets: an item costs several dependent round
        // trips (ticket, flag, statistics, the loads), which only other warps' items can hide - one CTA-wide worker
        // measured 12 us per 128-row item against the 2.2 us the HBM share of an SM allows.
        const int c8 = lane & 15;                          // 8 channels per lane, fixed: 32 lanes x 8 = 2 rows x 128
        const unsigned nch = (unsigned)((R + LR_CH - 1) / LR_CH);
        const unsigned per_utt = nch + 8;                  // 8 statistics parts, then the chunks
        const unsigned total = (unsigned)a.n_utt * per_utt;
        const uint64_t pol_s = lr_policy_evict_first();    // the residual stream passes through once: do not let it displace y
        float g[8], be[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { g[j] = __ldg(a.gamma + c8 * 8 + j); be[j] = __ldg(a.beta + c8 * 8 + j); }
        long ready_u = -1;                                 // utterances <= ready_u are known to be published
        float mean = 0.f, rstd = 0.f;
        long mr_u = -1;
        for (;;) {
            unsigned item = 0;
            if (lane == 0) item = atomicAdd(a.ticket1, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= total) break;
            const unsigned u = item / per_utt, c = item % per_utt;
            if (c < 8) {
                // statistics of utterance u with the arithmetic of row_stats_finalize_kernel (256 threads strided over the rows,
                // fp64, a tree per warp, then a tree over the 8 warp results): this item plays warp w = c of that block -
                // 8 items per utterance on 8 different warps, ~10 us each - and the one that finishes last combines the 8
                // results, so every partial sum and every tree has the same operands in the same order
                if (lane == 0) while (lr_ld_acquire(a.rows_done + u) < (unsigned)R) __nanosleep(200);
                __syncwarp();
                const float2* p = a.stats + (long)u * R;
                double s = 0.0, qd = 0.0;
                for (long rb = c * 32 + lane; rb < R; rb += 16 * 256) {       // 16 independent L2 loads in flight per lane
                    float2 v[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = rb + k * 256 < R ? __ldcg(p + rb + k * 256) : make_float2(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (rb + k * 256 < R) { s += (double)v[k].x; qd += (double)v[k].y; }
                }
                s = warp_sum(s); qd = warp_sum(qd);
                unsigned prior = 0;
                if (lane == 0) {
                    a.part8[(long)u * 16 + c] = s;
                    a.part8[(long)u * 16 + 8 + c] = qd;
                    __threadfence();
                    prior = atomicAdd(a.parts_done + u, 1u);
                }
                prior = __shfl_sync(0xffffffffu, prior, 0);
                if (prior == 7) {                                              // all 8 warp results are in: the final tree
                    __threadfence();
                    double rs = lane < 8 ? __ldcg(a.part8 + (long)u * 16 + lane) : 0.0;
                    double rq = lane < 8 ? __ldcg(a.part8 + (long)u * 16 + 8 + lane) : 0.0;
                    rs = warp_sum(rs); rq = warp_sum(rq);
                    if (lane == 0) {
                        const double cnt = (double)R * LR_N;
                        const double m = rs / cnt;
                        double var = rq / cnt - m * m;
                        if (var < 0.0) var = 0.0;
                        a.mean_rstd[2 * u] = (float)m;
                        a.mean_rstd[2 * u + 1] = (float)(1.0 / sqrt(var + a.eps));
                        __threadfence();
                        lr_st_release(a.flag + u, 1u);
                    }
                }
                continue;
            }
            if ((long)u > ready_u) {
                if (lane == 0) while (lr_ld_acquire(a.flag + u) == 0u) __nanosleep(100);
                __syncwarp();
                ready_u = u;
            }
            if ((long)u != mr_u) { mean = __ldcg(a.mean_rstd + 2 * u); rstd = __ldcg(a.mean_rstd + 2 * u + 1); mr_u = u; }
            const long r0 = (long)u * R + (long)(c - 8) * LR_CH;
            const long r1 = r0 + LR_CH < (long)(u + 1) * R ? r0 + LR_CH : (long)(u + 1) * R;
            const long end = r1 * 16;
            for (long base = r0 * 16; base < end; base += 256) {       // 16 rows = 256 uint4 per pass of the warp: 8 per lane
                uint4 yv[8], xv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {                  // all 16 loads of a lane are in flight before the first use
                    const long idx = base + i * 32 + lane;     // idx % 16 == c8 always (16 uint4 per row)
                    if (idx < end) {
                        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"        // L2 (written by another SM through TMA)
                                     : "=r"(yv[i].x), "=r"(yv[i].y), "=r"(yv[i].z), "=r"(yv[i].w) : "l"(a.y + idx));
                        asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                                     : "=r"(xv[i].x), "=r"(xv[i].y), "=r"(xv[i].z), "=r"(xv[i].w) : "l"(a.xb + idx), "l"(pol_s));
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long idx = base + i * 32 + lane;
                    if (idx < end) {
                        const uint32_t yw[4] = {yv[i].x, yv[i].y, yv[i].z, yv[i].w}, xw[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
                        uint32_t ob[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float2 yf = unpack_h16x2<kF16>(yw[h]), xf = unpack_h16x2<kF16>(xw[h]);
                            ob[h] = pack_h16x2<kF16>(norm_res1(xf.x, yf.x, mean, rstd, g[2 * h], be[2 * h]),
                                                     norm_res1(xf.y, yf.y, mean, rstd, g[2 * h + 1], be[2 * h + 1]));
                        }
                        asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                                     ::"l"(a.xb + idx), "r"(ob[0]), "r"(ob[1]), "r"(ob[2]), "r"(ob[3]), "l"(pol_s) : "memory");
                    }
                }
                if (a.discard) {
                    __syncwarp();                              // the 8 lanes that share a 128-byte line of y have used it
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const long idx = base + i * 32 + lane;
                        if (idx < end && (lane & 7) == 0) asm volatile("discard.global.L2 [%0], 128;" ::"l"(a.y + idx) : "memory");
                    }
                }
            }
            if (a.lead > 0) {
                __syncwarp();
                if (lane == 0) atomicAdd(a.chunks_done + u, 1u);       // flow control only: no data hangs on it
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem, 256);
}

template <int KB, bool kF16>
static int launch_lnr(const void* A, const void* W, void* C, LnrArgs args, cudaStream_t st) {
    constexpr int K = KB * 64;
    const int M = args.M;
    CUtensorMap tmA, tmW, tmC;
    const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {2, (uint64_t)K * 2};
    const uint32_t bA[2] = {64, 128};
    const uint64_t dW[2] = {(uint64_t)K, (uint64_t)LR_N}, sW[2] = {2, (uint64_t)K * 2};
    const uint32_t bW[2] = {64, (uint32_t)LR_N};
    const uint64_t dC[2] = {(uint64_t)LR_N, (uint64_t)M}, sC[2] = {2, (uint64_t)LR_N * 2};
    const uint32_t bC[2] = {64, 128};
    constexpr CUtensorMapDataType t16 = kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (make_tmap(&tmA, t16, 2, A, dA, sA, bA)) return 1;
    if (make_tmap(&tmW, t16, 2, W, dW, sW, bW)) return 1;
    if (make_tmap(&tmC, t16, 2, C, dC, sC, bC)) return 1;
    const size_t smem = (size_t)(KB + LR_AST + LR_CST) * LR_BLK + 1024;
    auto kern = linear_normres_kernel<KB, kF16>;
    DPRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    args.tiles = (int)cdiv(M, 128);
    kern<<<args.tiles < sms ? args.tiles : sms, LR_THREADS, smem, st>>>(tmA, tmW, tmC, args);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

}  // namespace dprnn

using namespace dprnn;

extern "C" size_t dprnn_linear_normres_workspace_bytes(int n_utt) {
    return 256 + (size_t)n_utt * 4 * sizeof(unsigned) + (size_t)n_utt * 16 * sizeof(double);     // chunks_done = 4th counter row
}

// xb[M,128] (16-bit, in place) += norm_u(h[M,K] @ W[128,K]^T + bias): Linear + GroupNorm(1,128) / gLN + residual of a
// half-block as ONE launch (see the header of this file).  y_scratch [M,128] 16-bit and stats_partial
// (dprnn_gemm_tc_stats_bytes(M)) are scratch; mean_rstd [M / rows_per_utt, 2] is an output.  Results are bit for bit those of
// dprnn_linear_h16out_stats followed by dprnn_norm_residual_h16res.
extern "C" int dprnn_linear_normres_h16(const void* h, const void* W, const float* bias, void* y_scratch, void* x_h16,
                                        const float* gamma, const float* beta, int M, int K, void* stats_partial,
                                        long rows_per_utt, float eps, float* mean_rstd, void* workspace, int discard_y,
                                        int h16, void* stream) {
    DPRNN_CHECK_ARG(h && W && bias && y_scratch && x_h16 && gamma && beta && stats_partial && mean_rstd && workspace);
    DPRNN_CHECK_ARG(M > 0 && (K == 128 || K == 256) && rows_per_utt >= 128 && M % rows_per_utt == 0);
    DPRNN_CHECK_ARG(h16 == DPRNN_H16_BF16 || h16 == DPRNN_H16_FP16);
    DPRNN_CHECK_ARG(((uintptr_t)h | (uintptr_t)W | (uintptr_t)y_scratch | (uintptr_t)x_h16 | (uintptr_t)workspace) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_utt = (int)(M / rows_per_utt);
    uint8_t* ws = (uint8_t*)workspace;
    DPRNN_CUDA(cudaMemsetAsync(ws, 0, 256 + (size_t)n_utt * 4 * sizeof(unsigned), st));
    LnrArgs a{};
    a.bias = bias; a.gamma = gamma; a.beta = beta;
    a.stats = (float2*)stats_partial;
    a.M = M; a.n_utt = n_utt; a.discard = discard_y & 1; a.rows_per_utt = rows_per_utt; a.eps = (double)eps;
    a.ticket = (unsigned*)ws; a.ticket1 = (unsigned*)(ws + 128);
    a.rows_done = (unsigned*)(ws + 256); a.flag = a.rows_done + n_utt; a.parts_done = a.flag + n_utt;
    a.chunks_done = a.parts_done + n_utt;
    a.lead = discard_y >> 8;                       // bits 8.. of the flag word: utterances pass 0 may run ahead (0 = unbounded)
    a.part8 = (double*)(ws + 256 + (size_t)n_utt * 4 * sizeof(unsigned));
    a.mean_rstd = mean_rstd;
    a.y = (const uint4*)y_scratch; a.xb = (uint4*)x_h16;
    if (h16 == DPRNN_H16_FP16)
        return K == 256 ? launch_lnr<4, true>(h, W, y_scratch, a, st) : launch_lnr<2, true>(h, W, y_scratch, a, st);
    return K == 256 ? launch_lnr<4, false>(h, W, y_scratch, a, st) : launch_lnr<2, false>(h, W, y_scratch, a, st);
}
