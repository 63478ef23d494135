// Ragged (variable-length) batches: the per-utterance stages of the path for utterances of different lengths
// packed back to back, so that a batch of full-length test utterances (cfg 3) gives exactly
// the per-utterance result of the reference's B = 1 loop (src/inferencers/inferencer_spe.py:25-45) - no padding
// enters any statistic, recurrence or softmax.
//
// Packed layouts (all channels-last):
//   frame space  : utterance b owns rows [frame_off[b], frame_off[b] + T_b) of which the first L_b = T_b - (ksz-1)
//                  are frames; the remaining ksz-1 rows are junk produced by running the encoder over the packed
//                  waveform (stride 1).  frame_utt[row] = b.  Waveform samples use the same indices.
//   chunk space  : utterance b owns chunks [chunk_off[b], chunk_off[b] + S_b), K rows each; chunk_utt[chunk] = b.
// Stages that are row-wise (1x1 convolutions, casts, BatchNorm-eval affine, the intra-chunk LSTM, the Linear) run on
// the packed buffers with the uniform kernels; this file holds the stages that need to know where an utterance ends.
#include "common.cuh"
#include <cuda_bf16.h>
#include "../../include/dprnn_b200.h"

namespace dprnn {

constexpr int kRaggedParts = 128;   // same partition as utt_stats_kernel: a packed batch gives bit-identical statistics

static inline unsigned rgrid(long total, int threads) {
    long g = (total + threads - 1) / threads;
    const long cap = 148L * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---------------------------------------------------------------- per-utterance statistics
__global__ void __launch_bounds__(256) utt_stats_ragged_kernel(const float* __restrict__ x, const long* __restrict__ off,
                                                               const long* __restrict__ len, int c4n,
                                                               double* __restrict__ partial) {
    __shared__ double scratch[32];
    const int b = blockIdx.y, p = blockIdx.x;
    const long elems4 = len[b] * c4n;
    const float4* xb = reinterpret_cast<const float4*>(x) + off[b] * c4n;
    const long per = (elems4 + gridDim.x - 1) / gridDim.x;
    const long beg = p * per, end = min(beg + per, elems4);
    double ds = 0.0, dq = 0.0;
    long i = beg + threadIdx.x;
    while (i < end) {
        float s = 0.f, q = 0.f;
        for (int u = 0; u < 16 && i < end; ++u, i += blockDim.x) {
            const float4 v = ld_stream(xb + i);
            s += (v.x + v.y) + (v.z + v.w);
            q = fmaf(v.x, v.x, q); q = fmaf(v.y, v.y, q); q = fmaf(v.z, v.z, q); q = fmaf(v.w, v.w, q);
        }
        ds += (double)s; dq += (double)q;
    }
    ds = block_sum(ds, scratch);
    dq = block_sum(dq, scratch);
    if (threadIdx.x == 0) {
        partial[((long)b * gridDim.x + p) * 2 + 0] = ds;
        partial[((long)b * gridDim.x + p) * 2 + 1] = dq;
    }
}

__global__ void utt_stats_ragged_finalize_kernel(const double* __restrict__ partial, const long* __restrict__ len, int C,
                                                 int nparts, double eps, float* __restrict__ mean_rstd) {
    const int b = blockIdx.x, lane = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (int p = lane; p < nparts; p += 32) {
        s += partial[((long)b * nparts + p) * 2 + 0];
        q += partial[((long)b * nparts + p) * 2 + 1];
    }
    s = warp_sum(s); q = warp_sum(q);
    if (lane == 0) {
        const double count = (double)len[b] * C;
        const double mean = s / count;
        double var = q / count - mean * mean;
        if (var < 0.0) var = 0.0;
        mean_rstd[2 * b + 0] = (float)mean;
        mean_rstd[2 * b + 1] = (float)(1.0 / sqrt(var + eps));
    }
}

// mean / rstd per utterance from the per-row {sum, sumsq} the tensor-core Linear emits (fixed reduction order)
__global__ void row_stats_finalize_ragged_kernel(const float2* __restrict__ partial, const long* __restrict__ row_off,
                                                 int cols, double eps, float* __restrict__ mean_rstd) {
    __shared__ double scratch[32];
    const int b = blockIdx.x;
    const long r0 = row_off[b], r1 = row_off[b + 1];
    double s = 0.0, q = 0.0;
    for (long r = r0 + threadIdx.x; r < r1; r += 256) {
        const float2 v = partial[r];
        s += (double)v.x; q += (double)v.y;
    }
    s = block_sum(s, scratch);
    q = block_sum(q, scratch);
    if (threadIdx.x == 0) {
        const double cnt = (double)(r1 - r0) * cols;
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        mean_rstd[2 * b] = (float)mean;
        mean_rstd[2 * b + 1] = (float)(1.0 / sqrt(var + eps));
    }
}

// ---------------------------------------------------------------- norm + residual on the chunk space
template <bool kYBf16, bool kF16>
__global__ void norm_residual_ragged_kernel(const void* __restrict__ yv, float* __restrict__ x,
                                            const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, const int* __restrict__ chunk_utt,
                                            long total4, long chunk4, int c4n, uint2* __restrict__ x_bf16) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const int b = __ldg(chunk_utt + idx / chunk4);
        const int c4 = (int)(idx % c4n);
        const float mean = __ldg(mean_rstd + 2 * b), rstd = __ldg(mean_rstd + 2 * b + 1);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        float4 v;
        if constexpr (kYBf16) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(yv) + idx);
            const float2 a = unpack_h16x2<kF16>(raw.x), c = unpack_h16x2<kF16>(raw.y);
            v = make_float4(a.x, a.y, c.x, c.y);
        } else {
            v = ld_stream(reinterpret_cast<const float4*>(yv) + idx);
        }
        float4 r = reinterpret_cast<float4*>(x)[idx];
        r.x += (v.x - mean) * rstd * g.x + be.x;
        r.y += (v.y - mean) * rstd * g.y + be.y;
        r.z += (v.z - mean) * rstd * g.z + be.z;
        r.w += (v.w - mean) * rstd * g.w + be.w;
        reinterpret_cast<float4*>(x)[idx] = r;
        if (x_bf16) x_bf16[idx] = make_uint2(pack_h16x2<kF16>(r.x, r.y), pack_h16x2<kF16>(r.z, r.w));
    }
}

// bf16-residual variant (Engine.residual_bf16; the uniform twin is norm_residual_bf16res_kernel in pointwise.cu - the
// per-element arithmetic is the same expression, so a packed batch stays bit-identical to per-utterance calls)
template <bool kF16>
__global__ void norm_residual_ragged_bf16res_kernel(const uint2* __restrict__ y, uint2* __restrict__ xb,
                                                    float* __restrict__ x_f32, const float* __restrict__ mean_rstd,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    const int* __restrict__ chunk_utt, long total4, long chunk4, int c4n) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const int b = __ldg(chunk_utt + idx / chunk4);
        const int c4 = (int)(idx % c4n);
        const float mean = __ldg(mean_rstd + 2 * b), rstd = __ldg(mean_rstd + 2 * b + 1);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        const uint2 yr = __ldg(y + idx);
        const uint2 xr = xb[idx];
        const float2 y01 = unpack_h16x2<kF16>(yr.x), y23 = unpack_h16x2<kF16>(yr.y);
        const float2 x01 = unpack_h16x2<kF16>(xr.x), x23 = unpack_h16x2<kF16>(xr.y);
        float4 r;
        r.x = x01.x + ((y01.x - mean) * rstd * g.x + be.x);
        r.y = x01.y + ((y01.y - mean) * rstd * g.y + be.y);
        r.z = x23.x + ((y23.x - mean) * rstd * g.z + be.z);
        r.w = x23.y + ((y23.y - mean) * rstd * g.w + be.w);
        if (x_f32) {
            reinterpret_cast<float4*>(x_f32)[idx] = r;
        } else {
            xb[idx] = make_uint2(pack_h16x2<kF16>(r.x, r.y), pack_h16x2<kF16>(r.z, r.w));
        }
    }
}

// ---------------------------------------------------------------- unfold / fold between frame and chunk space
template <bool kF16>
__global__ void unfold_ragged_kernel(const float* __restrict__ y, float* __restrict__ out,
                                     const int* __restrict__ chunk_utt, const long* __restrict__ chunk_off,
                                     const long* __restrict__ frame_off, const long* __restrict__ L, long total_chunks,
                                     int K, int P, int f4n, uint2* __restrict__ out_h16 = nullptr) {
    const long total = total_chunks * K * f4n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int f4 = (int)(idx % f4n);
        long r = idx / f4n;
        const int k = (int)(r % K);
        const long chunk = r / K;
        const int b = __ldg(chunk_utt + chunk);
        const long s = chunk - __ldg(chunk_off + b);
        const long t = s * P + k - K;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < __ldg(L + b)) v = __ldg(reinterpret_cast<const float4*>(y) + (__ldg(frame_off + b) + t) * f4n + f4);
        if (out) reinterpret_cast<float4*>(out)[idx] = v;
        if (out_h16)         // 16-bit copy for the first tensor-core LSTM layer (the rounding of dprnn_cast_h16)
            out_h16[idx] = make_uint2(pack_h16x2<kF16>(v.x, v.y), pack_h16x2<kF16>(v.z, v.w));
    }
}

__global__ void fold_prelu_ragged_kernel(const float* __restrict__ x, float* __restrict__ out,
                                         const int* __restrict__ frame_utt, const long* __restrict__ frame_off,
                                         const long* __restrict__ L, const long* __restrict__ chunk_off,
                                         const long* __restrict__ S, long total_rows, int K, int P, int f4n,
                                         const float* __restrict__ prelu_a) {
    const long total = total_rows * f4n;
    const float a = prelu_a ? __ldg(prelu_a) : 1.0f;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int f4 = (int)(idx % f4n);
        const long row = idx / f4n;
        const int b = __ldg(frame_utt + row);
        const long t = row - __ldg(frame_off + b);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < __ldg(L + b)) {
            const long Sb = __ldg(S + b), c0 = __ldg(chunk_off + b);
            long s_lo = t / P + 1, s_hi = (t + K) / P;
            if (s_hi > Sb - 1) s_hi = Sb - 1;
            for (long s = s_lo; s <= s_hi; ++s) {
                const long k = t + K - s * P;
                float4 v = ld_stream(reinterpret_cast<const float4*>(x) + ((c0 + s) * K + k) * f4n + f4);
                v.x = v.x >= 0.f ? v.x : a * v.x;
                v.y = v.y >= 0.f ? v.y : a * v.y;
                v.z = v.z >= 0.f ? v.z : a * v.z;
                v.w = v.w >= 0.f ? v.w : a * v.w;
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<float4*>(out)[idx] = acc;      // junk rows between utterances are zeroed
    }
}

// The last half-block's norm + residual (16-bit residual stream), PReLU and overlap-add in one pass (the ragged twin of
// norm_residual_fold_prelu_kernel in pointwise.cu): the per-element arithmetic is that of
// norm_residual_ragged_bf16res_kernel followed by fold_prelu_ragged_kernel, in the same order - bit-identical.
template <bool kF16>
__global__ void norm_residual_fold_prelu_ragged_kernel(const uint2* __restrict__ y, const uint2* __restrict__ xb,
                                                       const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ out,
                                                       const int* __restrict__ frame_utt, const long* __restrict__ frame_off,
                                                       const long* __restrict__ L, const long* __restrict__ chunk_off,
                                                       const long* __restrict__ S, long total_rows, int K, int P, int c4n,
                                                       const float* __restrict__ prelu_a) {
    const long total = total_rows * c4n;
    const float a = prelu_a ? __ldg(prelu_a) : 1.0f;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        const long row = idx / c4n;
        const int b = __ldg(frame_utt + row);
        const long t = row - __ldg(frame_off + b);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < __ldg(L + b)) {
            const float mean = __ldg(mean_rstd + 2 * b), rstd = __ldg(mean_rstd + 2 * b + 1);
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
            const long Sb = __ldg(S + b), c0 = __ldg(chunk_off + b);
            long s_lo = t / P + 1, s_hi = (t + K) / P;
            if (s_hi > Sb - 1) s_hi = Sb - 1;
            for (long s = s_lo; s <= s_hi; ++s) {
                const long k = t + K - s * P;
                const long e = ((c0 + s) * K + k) * c4n + c4;
                const uint2 yr = __ldg(y + e), xr = __ldg(xb + e);
                const float2 y01 = unpack_h16x2<kF16>(yr.x), y23 = unpack_h16x2<kF16>(yr.y);
                const float2 x01 = unpack_h16x2<kF16>(xr.x), x23 = unpack_h16x2<kF16>(xr.y);
                float4 v;
                v.x = x01.x + ((y01.x - mean) * rstd * g.x + be.x);
                v.y = x01.y + ((y01.y - mean) * rstd * g.y + be.y);
                v.z = x23.x + ((y23.x - mean) * rstd * g.z + be.z);
                v.w = x23.y + ((y23.y - mean) * rstd * g.w + be.w);
                v.x = v.x >= 0.f ? v.x : __fmul_rn(a, v.x);
                v.y = v.y >= 0.f ? v.y : __fmul_rn(a, v.y);
                v.z = v.z >= 0.f ? v.z : __fmul_rn(a, v.z);
                v.w = v.w >= 0.f ? v.w : __fmul_rn(a, v.w);
                acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
                acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
            }
        }
        reinterpret_cast<float4*>(out)[idx] = acc;      // junk rows between utterances are zeroed
    }
}

// ---------------------------------------------------------------- mask * enc -> ConvTranspose1d (stride 1)
__global__ void mask_decode_ragged_kernel(const float* __restrict__ mask, const float* __restrict__ enc,
                                          const float* __restrict__ wdec, float* __restrict__ out,
                                          const int* __restrict__ frame_utt, const long* __restrict__ frame_off,
                                          const long* __restrict__ L, long total_rows, int N, int ksz) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long o = warp; o < total_rows; o += nwarps) {
        const int b = __ldg(frame_utt + o);
        const long f0 = __ldg(frame_off + b), t = o - f0, Lb = __ldg(L + b);
        float acc = 0.f;
        for (int j = 0; j < ksz; ++j) {
            const long l = t - j;
            if (l < 0 || l >= Lb) continue;
            const float* m = mask + (f0 + l) * N;
            const float* e = enc + (f0 + l) * N;
            for (int c = lane; c < N; c += 32) acc = fmaf(m[c] * e[c], __ldg(wdec + c * ksz + j), acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[o] = acc;
    }
}

// kernel 2 / stride 1 / N = 64 fast path: same per-frame arithmetic as mask_decode_k2s1_kernel (pointwise.cu)
__global__ void __launch_bounds__(256) mask_decode_k2s1_ragged_kernel(const float* __restrict__ mask,
                                                                      const float* __restrict__ enc,
                                                                      const float* __restrict__ wdec,
                                                                      float* __restrict__ out,
                                                                      const int* __restrict__ frame_utt,
                                                                      const long* __restrict__ frame_off,
                                                                      const long* __restrict__ L, long total_rows) {
    __shared__ float sw[128];
    __shared__ float sp1[256];
    if (threadIdx.x < 128) sw[threadIdx.x] = wdec[threadIdx.x];
    const long row = (long)blockIdx.x * 256 + threadIdx.x;
    __syncthreads();
    long t = 0, Lb = 0;
    if (row < total_rows) {
        const int b = __ldg(frame_utt + row);
        t = row - __ldg(frame_off + b);
        Lb = __ldg(L + b);
    }
    float p0 = 0.f, p1 = 0.f;
    if (row < total_rows && t < Lb) decode_taps64(mask + row * 64, enc + row * 64, sw, p0, p1);
    sp1[threadIdx.x] = p1;
    float left = 0.f;
    if (threadIdx.x == 0 && row < total_rows && t >= 1) {
        float q0;
        decode_taps64(mask + (row - 1) * 64, enc + (row - 1) * 64, sw, q0, left);
    }
    __syncthreads();
    if (threadIdx.x > 0) left = (t >= 1) ? sp1[threadIdx.x - 1] : 0.f;
    if (row < total_rows) out[row] = p0 + left;          // t == 0: no left tap; t == L_b: only the left tap
}

// ---------------------------------------------------------------- attention fusion (dprnn_spe.py:212-229)
__global__ void att_scores_ragged_kernel(const float* __restrict__ enc, const float* __restrict__ s1,
                                         const float* __restrict__ s0, const float* __restrict__ wavg,
                                         const float* __restrict__ bavg, const float* __restrict__ v,
                                         float* __restrict__ score, const int* __restrict__ frame_utt,
                                         const long* __restrict__ frame_off, const long* __restrict__ La, long total_rows,
                                         int N, int ksz) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long o = warp; o < total_rows; o += nwarps) {
        const int b = __ldg(frame_utt + o);
        const long f0 = __ldg(frame_off + b), la = o - f0;
        if (la >= __ldg(La + b)) continue;              // scores of utterance b live at rows f0 .. f0 + La_b - 1
        float acc = 0.f;
        for (int c = lane; c < N; c += 32) {
            float avg = __ldg(bavg + c);
            for (int j = 0; j < ksz; ++j) {
                const float xn = enc[(f0 + la * ksz + j) * N + c] * s1[b * N + c] + s0[b * N + c];
                avg = fmaf(__ldg(wavg + c * ksz + j), xn, avg);
            }
            acc = fmaf(avg, v[b * N + c], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) score[o] = acc;
    }
}

__global__ void __launch_bounds__(256) softmax_ragged_kernel(float* __restrict__ x, const long* __restrict__ frame_off,
                                                             const long* __restrict__ La) {
    __shared__ float red[32];
    __shared__ double scratch[32];
    float* row = x + frame_off[blockIdx.x];
    const long n = La[blockIdx.x];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float m = -INFINITY;
    for (long i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, row[i]);
    m = warp_max(m);
    if (lane == 0) red[wid] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    double s = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) s += (double)expf(row[i] - m);
    s = block_sum(s, scratch);
    const float inv = (float)(1.0 / s);
    for (long i = threadIdx.x; i < n; i += blockDim.x) row[i] = expf(row[i] - m) * inv;
}

__global__ void att_rowscale_ragged_kernel(const float* __restrict__ sm, float* __restrict__ rowscale,
                                           const int* __restrict__ frame_utt, const long* __restrict__ frame_off,
                                           const long* __restrict__ L, const long* __restrict__ La, long total_rows) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total_rows; idx += (long)gridDim.x * blockDim.x) {
        const int b = __ldg(frame_utt + idx);
        const long f0 = __ldg(frame_off + b), l = idx - f0, Lb = __ldg(L + b), Lab = __ldg(La + b);
        float r = 1.0f;
        if (l < Lb) {
            long src;
            if (Lab == Lb) src = l;
            else if (Lb == 2 * Lab) src = l >> 1;
            else {
                const float scale = (float)Lab / (float)Lb;      // ATen nearest-upsample index rule
                src = (long)floorf(__fmul_rn((float)l, scale));
                if (src > Lab - 1) src = Lab - 1;
            }
            r = 1.0f + sm[f0 + src];
        }
        rowscale[idx] = r;
    }
}

// ---------------------------------------------------------------- speaker ResNet: pooling / time mean per utterance
__global__ void affine_add_prelu_pool3_ragged_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, const float* __restrict__ skip,
                                                     const float* __restrict__ prelu_a, float* __restrict__ out,
                                                     const int* __restrict__ out_utt, const long* __restrict__ in_off,
                                                     const long* __restrict__ out_off, long total_out_rows, int c4n) {
    const float a = __ldg(prelu_a);
    const long total = total_out_rows * c4n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        const long r = idx / c4n;
        const int b = __ldg(out_utt + r);
        const long lp = r - __ldg(out_off + b), i0 = __ldg(in_off + b);
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const long src = (i0 + lp * 3 + i) * c4n + c4;
            float4 v = reinterpret_cast<const float4*>(y)[src];
            const float4 k = reinterpret_cast<const float4*>(skip)[src];
            v.x = fmaf(v.x, sc.x, sh.x) + k.x; v.y = fmaf(v.y, sc.y, sh.y) + k.y;
            v.z = fmaf(v.z, sc.z, sh.z) + k.z; v.w = fmaf(v.w, sc.w, sh.w) + k.w;
            v.x = v.x >= 0.f ? v.x : a * v.x; v.y = v.y >= 0.f ? v.y : a * v.y;
            v.z = v.z >= 0.f ? v.z : a * v.z; v.w = v.w >= 0.f ? v.w : a * v.w;
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
        reinterpret_cast<float4*>(out)[idx] = m;
    }
}

__global__ void __launch_bounds__(256) time_sum_ragged_kernel(const float* __restrict__ x, float* __restrict__ emb,
                                                              const long* __restrict__ off, const long* __restrict__ len,
                                                              int C, const float* __restrict__ div) {
    __shared__ double sh[256];
    const int lanes = 256 / C;
    const int c = threadIdx.x % C, rl = threadIdx.x / C;
    const float* xb = x + off[blockIdx.x] * C;
    const long Lx = len[blockIdx.x];
    double s = 0.0;
    for (long l = rl; l < Lx; l += lanes) s += (double)xb[l * C + c];
    sh[threadIdx.x] = s;
    __syncthreads();
    if (rl == 0) {
        for (int i = 1; i < lanes; ++i) s += sh[i * C + c];
        emb[(long)blockIdx.x * C + c] = (float)s / div[blockIdx.x];
    }
}

}  // namespace dprnn

using namespace dprnn;

extern "C" {

size_t dprnn_utt_stats_ragged_workspace_bytes(int B) { return (size_t)B * kRaggedParts * 2 * sizeof(double); }

int dprnn_utt_stats_ragged(const float* x, int C, const long* off, const long* len, int B, float eps, void* workspace,
                           float* mean_rstd, void* stream) {
    DPRNN_CHECK_ARG(x && off && len && workspace && mean_rstd && B > 0 && B <= 65535 && C > 0 && C % 4 == 0);
    dim3 grid(kRaggedParts, B);
    utt_stats_ragged_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, off, len, C / 4, (double*)workspace);
    DPRNN_CHECK_LAUNCH();
    utt_stats_ragged_finalize_kernel<<<B, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, len, C, kRaggedParts,
                                                                        (double)eps, mean_rstd);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_row_stats_finalize_ragged(const void* stats_partial, const long* row_off, int B, int cols, float eps,
                                    float* mean_rstd, void* stream) {
    DPRNN_CHECK_ARG(stats_partial && row_off && mean_rstd && B > 0 && cols > 0);
    row_stats_finalize_ragged_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const float2*)stats_partial, row_off, cols,
                                                                         (double)eps, mean_rstd);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_norm_residual_ragged(const void* y, int y_is_bf16, float* x, const float* mean_rstd, const float* gamma,
                               const float* beta, const int* chunk_utt, long total_chunks, int K, int C, void* x_bf16,
                               void* stream) {
    DPRNN_CHECK_ARG(y && x && mean_rstd && gamma && beta && chunk_utt && total_chunks > 0 && K > 0 && C % 4 == 0);
    const long chunk4 = (long)K * (C / 4), total4 = total_chunks * chunk4;
    DPRNN_CHECK_ARG(y_is_bf16 >= 0 && y_is_bf16 <= 2);       // 0: y fp32; 1: y and x_bf16 bf16; 2: y and x_bf16 fp16
    auto kern = y_is_bf16 == 2 ? norm_residual_ragged_kernel<true, true>
                               : (y_is_bf16 ? norm_residual_ragged_kernel<true, false> : norm_residual_ragged_kernel<false, false>);
    kern<<<rgrid(total4, 256), 256, 0, (cudaStream_t)stream>>>(y, x, mean_rstd, gamma, beta, chunk_utt, total4, chunk4,
                                                               C / 4, (uint2*)x_bf16);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_norm_residual_ragged_h16res(const void* y_h16, void* x_h16, float* x_f32_out, const float* mean_rstd,
                                      const float* gamma, const float* beta, const int* chunk_utt, long total_chunks,
                                      int K, int C, int h16, void* stream) {
    DPRNN_CHECK_ARG(y_h16 && x_h16 && mean_rstd && gamma && beta && chunk_utt && total_chunks > 0 && K > 0 && C % 4 == 0);
    DPRNN_CHECK_ARG(h16 == DPRNN_H16_BF16 || h16 == DPRNN_H16_FP16);
    const long chunk4 = (long)K * (C / 4), total4 = total_chunks * chunk4;
    auto kern = h16 ? norm_residual_ragged_bf16res_kernel<true> : norm_residual_ragged_bf16res_kernel<false>;
    kern<<<rgrid(total4, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint2*)y_h16, (uint2*)x_h16, x_f32_out, mean_rstd, gamma, beta, chunk_utt, total4, chunk4, C / 4);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
int dprnn_norm_residual_ragged_bf16res(const void* y_bf16, void* x_bf16, float* x_f32_out, const float* mean_rstd,
                                       const float* gamma, const float* beta, const int* chunk_utt, long total_chunks,
                                       int K, int C, void* stream) {
    return dprnn_norm_residual_ragged_h16res(y_bf16, x_bf16, x_f32_out, mean_rstd, gamma, beta, chunk_utt, total_chunks,
                                             K, C, DPRNN_H16_BF16, stream);
}

int dprnn_unfold_ragged(const float* y, float* x, const int* chunk_utt, const long* chunk_off, const long* frame_off,
                        const long* L, long total_chunks, int K, int P, int F, void* stream) {
    DPRNN_CHECK_ARG(y && x && chunk_utt && chunk_off && frame_off && L && total_chunks > 0 && K > 0 && P > 0 && F % 4 == 0);
    unfold_ragged_kernel<false><<<rgrid(total_chunks * K * (F / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        y, x, chunk_utt, chunk_off, frame_off, L, total_chunks, K, P, F / 4);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_unfold_ragged_h16(const float* y, float* x, void* x_h16, const int* chunk_utt, const long* chunk_off,
                            const long* frame_off, const long* L, long total_chunks, int K, int P, int F, int h16,
                            void* stream) {
    DPRNN_CHECK_ARG(y && x_h16 && chunk_utt && chunk_off && frame_off && L && total_chunks > 0 && K > 0 && P > 0 && F % 4 == 0);
    DPRNN_CHECK_ARG(h16 == DPRNN_H16_BF16 || h16 == DPRNN_H16_FP16);
    auto kern = h16 ? unfold_ragged_kernel<true> : unfold_ragged_kernel<false>;
    kern<<<rgrid(total_chunks * K * (F / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        y, x, chunk_utt, chunk_off, frame_off, L, total_chunks, K, P, F / 4, (uint2*)x_h16);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_norm_residual_fold_prelu_ragged_h16(const void* y_h16, const void* x_h16, const float* mean_rstd,
                                              const float* gamma, const float* beta, float* out, const int* frame_utt,
                                              const long* frame_off, const long* L, const long* chunk_off, const long* S,
                                              long total_rows, int K, int P, int F, const float* prelu_a, int h16,
                                              void* stream) {
    DPRNN_CHECK_ARG(y_h16 && x_h16 && mean_rstd && gamma && beta && out && frame_utt && frame_off && L && chunk_off && S);
    DPRNN_CHECK_ARG(total_rows > 0 && K > 0 && P > 0 && F % 4 == 0);
    DPRNN_CHECK_ARG(h16 == DPRNN_H16_BF16 || h16 == DPRNN_H16_FP16);
    auto kern = h16 ? norm_residual_fold_prelu_ragged_kernel<true> : norm_residual_fold_prelu_ragged_kernel<false>;
    kern<<<rgrid(total_rows * (F / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint2*)y_h16, (const uint2*)x_h16, mean_rstd, gamma, beta, out, frame_utt, frame_off, L, chunk_off, S,
        total_rows, K, P, F / 4, prelu_a);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_fold_prelu_ragged(const float* x, float* out, const int* frame_utt, const long* frame_off, const long* L,
                            const long* chunk_off, const long* S, long total_rows, int K, int P, int F,
                            const float* prelu_a, void* stream) {
    DPRNN_CHECK_ARG(x && out && frame_utt && frame_off && L && chunk_off && S && total_rows > 0 && K > 0 && P > 0 && F % 4 == 0);
    fold_prelu_ragged_kernel<<<rgrid(total_rows * (F / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        x, out, frame_utt, frame_off, L, chunk_off, S, total_rows, K, P, F / 4, prelu_a);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_mask_decode_ragged(const float* mask, const float* enc, const float* wdec, float* out, const int* frame_utt,
                             const long* frame_off, const long* L, long total_rows, int N, int ksz, void* stream) {
    DPRNN_CHECK_ARG(mask && enc && wdec && out && frame_utt && frame_off && L && total_rows > 0 && N > 0 && ksz > 0);
    if (ksz == 2 && N == 64 && ((uintptr_t)mask | (uintptr_t)enc) % 16 == 0) {
        mask_decode_k2s1_ragged_kernel<<<cdiv(total_rows, 256), 256, 0, (cudaStream_t)stream>>>(
            mask, enc, wdec, out, frame_utt, frame_off, L, total_rows);
        DPRNN_CHECK_LAUNCH();
        return 0;
    }
    mask_decode_ragged_kernel<<<rgrid(total_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        mask, enc, wdec, out, frame_utt, frame_off, L, total_rows, N, ksz);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_att_rowscale_ragged(const float* enc, const float* s1, const float* s0, const float* wavg, const float* bavg,
                              const float* v, float* scores, float* rowscale, const int* frame_utt,
                              const long* frame_off, const long* L, const long* La, int B, long total_rows, int N,
                              int ksz, void* stream) {
    DPRNN_CHECK_ARG(enc && s1 && s0 && wavg && bavg && v && scores && rowscale && frame_utt && frame_off && L && La);
    DPRNN_CHECK_ARG(B > 0 && total_rows > 0 && N > 0 && ksz > 0);
    cudaStream_t st = (cudaStream_t)stream;
    att_scores_ragged_kernel<<<rgrid(total_rows * 32, 256), 256, 0, st>>>(enc, s1, s0, wavg, bavg, v, scores, frame_utt,
                                                                          frame_off, La, total_rows, N, ksz);
    DPRNN_CHECK_LAUNCH();
    softmax_ragged_kernel<<<B, 256, 0, st>>>(scores, frame_off, La);
    DPRNN_CHECK_LAUNCH();
    att_rowscale_ragged_kernel<<<rgrid(total_rows, 256), 256, 0, st>>>(scores, rowscale, frame_utt, frame_off, L, La,
                                                                       total_rows);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_affine_add_prelu_pool3_ragged(const float* y, const float* scale, const float* shift, const float* skip,
                                        const float* prelu_a, float* out, const int* out_utt, const long* in_off,
                                        const long* out_off, long total_out_rows, int C, void* stream) {
    DPRNN_CHECK_ARG(y && scale && shift && skip && prelu_a && out && out_utt && in_off && out_off && C % 4 == 0);
    DPRNN_CHECK_ARG(total_out_rows > 0);
    affine_add_prelu_pool3_ragged_kernel<<<rgrid(total_out_rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        y, scale, shift, skip, prelu_a, out, out_utt, in_off, out_off, total_out_rows, C / 4);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_time_sum_ragged(const float* x, float* emb, const long* off, const long* len, int B, int C, const float* div,
                          void* stream) {
    DPRNN_CHECK_ARG(x && emb && off && len && div && B > 0 && C > 0 && 256 % C == 0);
    time_sum_ragged_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, emb, off, len, C, div);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
