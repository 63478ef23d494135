// Memory-bound stages of the DPRNN forward: waveform encoder / decoder, chunk unfold / fold,
// per-utterance norm statistics + apply, speaker-fusion helpers, BatchNorm / PReLU / MaxPool of the
// speaker ResNet.  All activations are fp32, channels-last ([B, L, C] / [B, S, K, F]).
// The reference line each entry point replaces is cited in include/dprnn_b200.h.
#include "common.cuh"
#include <cuda_bf16.h>
#include "../../include/dprnn_b200.h"

namespace dprnn {

constexpr int kStatParts = 128;   // partial sums per utterance (deterministic two-level reduction)

// ------------------------------------------------------------------------------------------
// encoder: enc[b,l,c] = relu(sum_j w[c,j] * x[b, l*stride + j])
// ------------------------------------------------------------------------------------------
__global__ void encoder_kernel(const float* __restrict__ wave, const float* __restrict__ w,
                               float* __restrict__ enc, int B, int T, long L, int N, int ksz, int stride) {
    const int n4 = N >> 2;
    const long total = (long)B * L * n4;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % n4);
        const long row = idx / n4;
        const long b = row / L, l = row % L;
        const float* x = wave + b * T + l * stride;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int j = 0; j < ksz; ++j) {
            const float xv = __ldg(x + j);
            const float* wc = w + (long)(c4 * 4) * ksz + j;
            a0 = fmaf(__ldg(wc), xv, a0);
            a1 = fmaf(__ldg(wc + ksz), xv, a1);
            a2 = fmaf(__ldg(wc + 2 * ksz), xv, a2);
            a3 = fmaf(__ldg(wc + 3 * ksz), xv, a3);
        }
        float4 o = make_float4(fmaxf(a0, 0.f), fmaxf(a1, 0.f), fmaxf(a2, 0.f), fmaxf(a3, 0.f));
        reinterpret_cast<float4*>(enc)[idx] = o;
    }
}

// kernel 2, stride 1 (every shipped config): the thread's 4 x 2 weights stay in registers (its channel group is fixed: the
// grid stride is a multiple of n4), 32-bit index arithmetic (the launcher checks that the indices fit) - the generic
// kernel spends its time in 64-bit divisions and weight re-loads (2.2 TB/s).  Same fmaf chain, same bits.
__global__ void __launch_bounds__(256) encoder_k2s1_kernel(const float* __restrict__ wave, const float* __restrict__ w,
                                                           float4* __restrict__ enc, unsigned B, unsigned T, unsigned L,
                                                           unsigned n4) {
    const unsigned c4 = threadIdx.x % n4;
    const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * c4), wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * c4 + 1);
    // w[c][j], c = 4 c4 + {0,1,2,3}: wa = {w00, w01, w10, w11}, wb = {w20, w21, w30, w31}
    const unsigned total = B * L * n4;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned row = idx / n4;
        const unsigned b = row / L, l = row - b * L;
        const float* x = wave + (size_t)b * T + l;
        const float x0 = __ldg(x), x1 = __ldg(x + 1);
        const float a0 = fmaf(wa.y, x1, fmaf(wa.x, x0, 0.f)), a1 = fmaf(wa.w, x1, fmaf(wa.z, x0, 0.f));
        const float a2 = fmaf(wb.y, x1, fmaf(wb.x, x0, 0.f)), a3 = fmaf(wb.w, x1, fmaf(wb.z, x0, 0.f));
        enc[idx] = make_float4(fmaxf(a0, 0.f), fmaxf(a1, 0.f), fmaxf(a2, 0.f), fmaxf(a3, 0.f));
    }
}

// ------------------------------------------------------------------------------------------
// per-utterance statistics over a contiguous slab: partial[b][p] = {sum, sumsq} in fp64
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) utt_stats_kernel(const float* __restrict__ x, double* __restrict__ partial,
                                                        long elems4) {
    __shared__ double scratch[32];
    const int b = blockIdx.y, p = blockIdx.x;
    const float4* xb = reinterpret_cast<const float4*>(x) + (long)b * elems4;
    const long per = (elems4 + gridDim.x - 1) / gridDim.x;
    const long beg = p * per, end = min(beg + per, elems4);
    double ds = 0.0, dq = 0.0;
    long i = beg + threadIdx.x;
    while (i < end) {
        float s = 0.f, q = 0.f;
#pragma unroll 4
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

__global__ void utt_stats_finalize_kernel(const double* __restrict__ partial, float* __restrict__ mean_rstd,
                                          int nparts, double count, double eps) {
    const int b = blockIdx.x, lane = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (int p = lane; p < nparts; p += 32) {
        s += partial[((long)b * nparts + p) * 2 + 0];
        q += partial[((long)b * nparts + p) * 2 + 1];
    }
    s = warp_sum(s); q = warp_sum(q);
    if (lane == 0) {
        const double mean = s / count;
        double var = q / count - mean * mean;
        if (var < 0.0) var = 0.0;
        mean_rstd[2 * b + 0] = (float)mean;
        mean_rstd[2 * b + 1] = (float)(1.0 / sqrt(var + eps));
    }
}

// s1[b,c] = gamma*rstd*mul ; s0[b,c] = (beta - mean*gamma*rstd)*mul
__global__ void norm_affine_kernel(const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ mulc,
                                   float* __restrict__ s1, float* __restrict__ s0, int B, int C) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * C) return;
    const int b = idx / C, c = idx % C;
    const float mean = mean_rstd[2 * b], rstd = mean_rstd[2 * b + 1];
    const float m = mulc ? mulc[idx] : 1.0f;
    const float g = gamma[c] * rstd;
    s1[idx] = g * m;
    s0[idx] = (beta[c] - mean * g) * m;
}

// x[b,r,c] += (y[b,r,c] - mean_b) * rstd_b * gamma_c + beta_c
template <bool kF16>
__global__ void norm_residual_kernel(const float* __restrict__ y, float* __restrict__ x,
                                     const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, long total4, long per_utt4, int c4n,
                                     uint2* __restrict__ x_bf16, float* __restrict__ x_out = nullptr) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / per_utt4;
        const int c4 = (int)(idx % c4n);
        const float mean = __ldg(mean_rstd + 2 * b), rstd = __ldg(mean_rstd + 2 * b + 1);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        const float4 v = ld_stream(reinterpret_cast<const float4*>(y) + idx);
        float4 r = reinterpret_cast<float4*>(x)[idx];
        r.x += (v.x - mean) * rstd * g.x + be.x;
        r.y += (v.y - mean) * rstd * g.y + be.y;
        r.z += (v.z - mean) * rstd * g.z + be.z;
        r.w += (v.w - mean) * rstd * g.w + be.w;
        reinterpret_cast<float4*>(x_out ? x_out : x)[idx] = r;      // x_out: out of place (x is left as it was)
        if (x_bf16)      // 16-bit shadow copy: the TMA-fed A operand of the next tensor-core LSTM layer
            x_bf16[idx] = make_uint2(pack_h16x2<kF16>(r.x, r.y), pack_h16x2<kF16>(r.z, r.w));
    }
}

// Same with y in bf16 / fp16 (output of the tensor-core Linear), 8 elements per thread
template <bool kF16>
__global__ void norm_residual_ybf16_kernel(const uint4* __restrict__ y, float* __restrict__ x,
                                           const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                           const float* __restrict__ beta, long total8, long per_utt8, int c8n,
                                           uint4* __restrict__ x_bf16) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total8; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / per_utt8;
        const int c8 = (int)(idx % c8n);
        const float mean = __ldg(mean_rstd + 2 * b), rstd = __ldg(mean_rstd + 2 * b + 1);
        uint4 yv;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(yv.x), "=r"(yv.y), "=r"(yv.z), "=r"(yv.w) : "l"(y + idx));
        const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
        float4 r[2] = {reinterpret_cast<float4*>(x)[2 * idx], reinterpret_cast<float4*>(x)[2 * idx + 1]};
        uint32_t ob[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c8 + h);
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c8 + h);
            const float2 y01 = unpack_h16x2<kF16>(yw[2 * h]), y23 = unpack_h16x2<kF16>(yw[2 * h + 1]);
            r[h].x += (y01.x - mean) * rstd * g.x + be.x;
            r[h].y += (y01.y - mean) * rstd * g.y + be.y;
            r[h].z += (y23.x - mean) * rstd * g.z + be.z;
            r[h].w += (y23.y - mean) * rstd * g.w + be.w;
            ob[2 * h] = pack_h16x2<kF16>(r[h].x, r[h].y);
            ob[2 * h + 1] = pack_h16x2<kF16>(r[h].z, r[h].w);
        }
        reinterpret_cast<float4*>(x)[2 * idx] = r[0];
        reinterpret_cast<float4*>(x)[2 * idx + 1] = r[1];
        if (x_bf16) x_bf16[idx] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
    }
}

// Opt-in "bf16 residual stream" variant of the above (Engine.residual_bf16): the residual x lives only in its bf16 copy,
// x_bf16 <- bf16(float(x_bf16) + norm(y)); the fp32 master is written once, by the last half-block (x_f32 != NULL), for
// the fold.  2.4 GB instead of 4.77 GB per launch at B = 64; costs ~8 dB of the bf16 mode's 52 dB agreement with the
// reference, hence not the default.
template <bool kF16>
__global__ void norm_residual_bf16res_kernel(const uint4* __restrict__ y, uint4* __restrict__ xb, float* __restrict__ x_f32,
                                             const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, long total8, long per_utt8, int c8n) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total8; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / per_utt8;
        const int c8 = (int)(idx % c8n);
        const float mean = __ldg(mean_rstd + 2 * b), rstd = __ldg(mean_rstd + 2 * b + 1);
        uint4 yv;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(yv.x), "=r"(yv.y), "=r"(yv.z), "=r"(yv.w) : "l"(y + idx));
        const uint4 xv = xb[idx];
        const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w}, xw[4] = {xv.x, xv.y, xv.z, xv.w};
        float4 r[2];
        uint32_t ob[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c8 + h);
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c8 + h);
            const float2 y01 = unpack_h16x2<kF16>(yw[2 * h]), y23 = unpack_h16x2<kF16>(yw[2 * h + 1]);
            const float2 x01 = unpack_h16x2<kF16>(xw[2 * h]), x23 = unpack_h16x2<kF16>(xw[2 * h + 1]);
            r[h].x = norm_res1(x01.x, y01.x, mean, rstd, g.x, be.x);
            r[h].y = norm_res1(x01.y, y01.y, mean, rstd, g.y, be.y);
            r[h].z = norm_res1(x23.x, y23.x, mean, rstd, g.z, be.z);
            r[h].w = norm_res1(x23.y, y23.y, mean, rstd, g.w, be.w);
            ob[2 * h] = pack_h16x2<kF16>(r[h].x, r[h].y);
            ob[2 * h + 1] = pack_h16x2<kF16>(r[h].z, r[h].w);
        }
        if (x_f32) {
            reinterpret_cast<float4*>(x_f32)[2 * idx] = r[0];
            reinterpret_cast<float4*>(x_f32)[2 * idx + 1] = r[1];
        } else {
            xb[idx] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
        }
    }
}

// out = (a*p_scale[b,c] + p_shift[b,c]) * rowscale[row] + p_add[b,c]: the per-utterance norm + speaker fusion that the
// exact-fp32 GEMM applies as a prologue, materialised for the tensor-core GEMM (whose operands go smem -> MMA directly)
__global__ void prologue_apply_kernel(const float* __restrict__ a, float* __restrict__ out, long total4, int c4n,
                                      long rows_per_utt, const float* __restrict__ p_scale,
                                      const float* __restrict__ p_shift, const float* __restrict__ p_add,
                                      const float* __restrict__ rowscale, const int* __restrict__ row_utt) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        const long row = idx / c4n, b = row_utt ? (long)__ldg(row_utt + row) : row / rows_per_utt;
        float4 v = ld_stream(reinterpret_cast<const float4*>(a) + idx);
        if (p_scale) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p_scale) + b * c4n + c4);
            const float4 sh = __ldg(reinterpret_cast<const float4*>(p_shift) + b * c4n + c4);
            v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        }
        if (rowscale) { const float rs = __ldg(rowscale + row); v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs; }
        if (p_add) {
            const float4 ad = __ldg(reinterpret_cast<const float4*>(p_add) + b * c4n + c4);
            v.x += ad.x; v.y += ad.y; v.z += ad.z; v.w += ad.w;
        }
        reinterpret_cast<float4*>(out)[idx] = v;
    }
}

template <bool kF16>
__global__ void cast_bf16_kernel(const float* __restrict__ x, uint2* __restrict__ out, long total4) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const float4 r = reinterpret_cast<const float4*>(x)[idx];
        out[idx] = make_uint2(pack_h16x2<kF16>(r.x, r.y), pack_h16x2<kF16>(r.z, r.w));
    }
}

// ------------------------------------------------------------------------------------------
// unfold: out[b,s,k,:] = y[b, s*P + k - K, :] (zero outside [0,L))
// ------------------------------------------------------------------------------------------
template <bool kF16>
__global__ void unfold_kernel(const float* __restrict__ y, float* __restrict__ out, int B, long L, int S, int K,
                              int P, int f4n, uint2* __restrict__ out_bf16) {
    const long total = (long)B * S * K * f4n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int f4 = (int)(idx % f4n);
        long r = idx / f4n;
        const int k = (int)(r % K); r /= K;
        const int s = (int)(r % S);
        const long b = r / S;
        const long t = (long)s * P + k - K;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < L) v = __ldg(reinterpret_cast<const float4*>(y) + (b * L + t) * f4n + f4);
        if (out) reinterpret_cast<float4*>(out)[idx] = v;      // out == NULL: only the 16-bit copy is wanted
        if (out_bf16)        // 16-bit shadow for the first tensor-core LSTM layer (saves a separate cast pass)
            out_bf16[idx] = make_uint2(pack_h16x2<kF16>(v.x, v.y), pack_h16x2<kF16>(v.z, v.w));
    }
}

// unfold with the 16-bit output only (the default of the tensor-core modes): 8 channels per thread and 32-bit index
// arithmetic (the launcher checks that every index fits) - the generic kernel above spends its time in 64-bit divisions
// once it no longer writes the fp32 copy.  Same values, same rounding.
template <bool kF16>
__global__ void unfold_h16_rows_kernel(const float4* __restrict__ y, uint4* __restrict__ out16, unsigned rows, unsigned L,
                                       unsigned S, unsigned K, unsigned P, unsigned c8n) {
    const unsigned total = rows * c8n;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned c8 = idx % c8n, row = idx / c8n;
        const unsigned k = row % K, r2 = row / K;
        const unsigned s = r2 % S, b = r2 / S;
        const int t = (int)(s * P + k) - (int)K;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (t >= 0 && (unsigned)t < L) {
            const float4* src = y + ((size_t)b * L + (unsigned)t) * (2 * c8n) + 2 * c8;
            v0 = __ldg(src); v1 = __ldg(src + 1);
        }
        out16[idx] = make_uint4(pack_h16x2<kF16>(v0.x, v0.y), pack_h16x2<kF16>(v0.z, v0.w),
                                pack_h16x2<kF16>(v1.x, v1.y), pack_h16x2<kF16>(v1.z, v1.w));
    }
}

// fold (+ optional PReLU on the way in): out[b,t,:] = sum_{s: 0 <= t+K-sP < K} prelu(x[b,s,t+K-sP,:])
__global__ void fold_prelu_kernel(const float* __restrict__ x, float* __restrict__ out, int B, long L, int S,
                                  int K, int P, int f4n, const float* __restrict__ prelu_a) {
    const long total = (long)B * L * f4n;
    const float a = prelu_a ? __ldg(prelu_a) : 1.0f;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int f4 = (int)(idx % f4n);
        const long r = idx / f4n;
        const long t = r % L, b = r / L;
        long s_lo = t / P + 1, s_hi = (t + K) / P;
        if (s_hi > S - 1) s_hi = S - 1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long s = s_lo; s <= s_hi; ++s) {
            const long k = t + K - s * P;
            float4 v = ld_stream(reinterpret_cast<const float4*>(x) + ((b * S + s) * K + k) * f4n + f4);
            v.x = v.x >= 0.f ? v.x : a * v.x;
            v.y = v.y >= 0.f ? v.y : a * v.y;
            v.z = v.z >= 0.f ? v.z : a * v.z;
            v.w = v.w >= 0.f ? v.w : a * v.w;
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[idx] = acc;
    }
}

// The last half-block's norm + residual (16-bit residual stream), the PReLU and the fold in ONE pass: every (s, k) position
// of the last Linear output feeds exactly one output frame, so x + norm(y) never has to exist as an fp32 [B,S,K,F] tensor
// (written by norm_residual_bf16res_kernel and read back by fold_prelu_kernel: 2 x 1.59 GB at B = 64).  Same arithmetic
// in the same order as those two kernels (norm_res1, predicated multiply, sum over s ascending): bit-identical.
// I = the integer type of the index arithmetic: unsigned when every index fits 32 bits (the launcher checks), else long -
// the divisions are most of the kernel's instructions.
template <bool kF16, typename I>
__global__ void norm_residual_fold_prelu_kernel(const uint4* __restrict__ y, const uint4* __restrict__ xb,
                                                const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, float* __restrict__ out, int B_, long L_,
                                                int S_, int K_, int P_, int c8n_, const float* __restrict__ prelu_a) {
    const I L = (I)L_, S = (I)S_, K = (I)K_, P = (I)P_, c8n = (I)c8n_;
    const I total = (I)B_ * L * c8n;
    const float a = prelu_a ? __ldg(prelu_a) : 1.0f;
    for (I idx = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; idx < total; idx += (I)gridDim.x * (I)blockDim.x) {
        const I c8 = idx % c8n;
        const I r = idx / c8n;
        const I t = r % L, b = r / L;
        const float mean = __ldg(mean_rstd + 2 * b), rstd = __ldg(mean_rstd + 2 * b + 1);
        I s_lo = t / P + 1, s_hi = (t + K) / P;
        if (s_hi > S - 1) s_hi = S - 1;
        float g[8], be[8], acc[8];
        {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c8), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c8 + 1);
            const float4 e0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c8), e1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c8 + 1);
            g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
            be[0] = e0.x; be[1] = e0.y; be[2] = e0.z; be[3] = e0.w; be[4] = e1.x; be[5] = e1.y; be[6] = e1.z; be[7] = e1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (I s = s_lo; s <= s_hi; ++s) {
            const I k = t + K - s * P;
            const size_t e = (size_t)(((b * S + s) * K + k) * c8n + c8);
            uint4 yv;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(yv.x), "=r"(yv.y), "=r"(yv.z), "=r"(yv.w) : "l"(y + e));
            const uint4 xv = __ldg(xb + e);
            const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w}, xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 yf = unpack_h16x2<kF16>(yw[q]), xf = unpack_h16x2<kF16>(xw[q]);
                float v0 = norm_res1(xf.x, yf.x, mean, rstd, g[2 * q], be[2 * q]);
                float v1 = norm_res1(xf.y, yf.y, mean, rstd, g[2 * q + 1], be[2 * q + 1]);
                v0 = v0 >= 0.f ? v0 : __fmul_rn(a, v0);
                v1 = v1 >= 0.f ? v1 : __fmul_rn(a, v1);
                acc[2 * q] = __fadd_rn(acc[2 * q], v0);
                acc[2 * q + 1] = __fadd_rn(acc[2 * q + 1], v1);
            }
        }
        float4* o = reinterpret_cast<float4*>(out) + (size_t)idx * 2;
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// ------------------------------------------------------------------------------------------
// mask * enc -> ConvTranspose1d(N -> 1, ksz, stride): one warp per output sample (general ksz / stride)
// ------------------------------------------------------------------------------------------
__global__ void mask_decode_kernel(const float* __restrict__ mask, long mask_utt_stride,
                                   const float* __restrict__ enc, const float* __restrict__ wdec,
                                   float* __restrict__ out, long out_utt_stride, int B, long L, int T, int N,
                                   int ksz, int stride) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long o = warp; o < (long)B * T; o += nwarps) {
        const long b = o / T, t = o % T;
        float acc = 0.f;
        for (int j = 0; j < ksz; ++j) {
            const long tl = t - j;
            if (tl < 0 || tl % stride) continue;
            const long l = tl / stride;
            if (l >= L) continue;
            const float* m = mask + b * mask_utt_stride + l * N;
            const float* e = enc + (b * L + l) * N;
            for (int c = lane; c < N; c += 32) acc = fmaf(m[c] * e[c], __ldg(wdec + c * ksz + j), acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[b * out_utt_stride + t] = acc;
    }
}

// The shipped geometry (kernel 2, stride 1, N = 64): one thread per frame reads its 2 x 256 contiguous bytes with
// sixteen 16-byte loads in flight, forms the two taps p0 = <m.e, w[:,0]>, p1 = <m.e, w[:,1]> and
// out[t] = p0[t] + p1[t-1]; the neighbour's p1 comes through shared memory (thread 0 recomputes its left halo).
__global__ void __launch_bounds__(256) mask_decode_k2s1_kernel(const float* __restrict__ mask, long mask_utt_stride,
                                                               const float* __restrict__ enc,
                                                               const float* __restrict__ wdec, float* __restrict__ out,
                                                               long out_utt_stride, long L, int blocks_per_utt) {
    __shared__ float sw[128];
    __shared__ float sp1[256];
    if (threadIdx.x < 128) sw[threadIdx.x] = wdec[threadIdx.x];         // wdec [64, 2]
    const long b = blockIdx.x / blocks_per_utt;
    const long t = (long)(blockIdx.x % blocks_per_utt) * 256 + threadIdx.x;   // output sample, 0 .. L
    __syncthreads();
    const float* mb = mask + b * mask_utt_stride;
    const float* eb = enc + b * L * 64;
    float p0 = 0.f, p1 = 0.f;
    if (t < L) decode_taps64(mb + t * 64, eb + t * 64, sw, p0, p1);
    sp1[threadIdx.x] = p1;
    float left = 0.f;
    if (threadIdx.x == 0 && t >= 1 && t - 1 < L) {
        float q0;
        decode_taps64(mb + (t - 1) * 64, eb + (t - 1) * 64, sw, q0, left);
    }
    __syncthreads();
    if (threadIdx.x > 0) left = sp1[threadIdx.x - 1];
    if (t <= L) out[b * out_utt_stride + t] = p0 + left;
}

// out = mask * enc (the IRA re-embedding input d0, dprnn_spe_ira.py:79-80)
__global__ void mask_apply_kernel(const float* __restrict__ mask, const float* __restrict__ enc,
                                  float* __restrict__ out, long total4) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const float4 m = reinterpret_cast<const float4*>(mask)[idx];
        const float4 e = reinterpret_cast<const float4*>(enc)[idx];
        reinterpret_cast<float4*>(out)[idx] = make_float4(m.x * e.x, m.y * e.y, m.z * e.z, m.w * e.w);
    }
}

// ------------------------------------------------------------------------------------------
// attention fusion helpers
// ------------------------------------------------------------------------------------------
// score[b,la] = sum_c v[b,c] * (bavg[c] + sum_j wavg[c,j] * (enc[b,la*k+j,c]*s1[b,c] + s0[b,c]))
__global__ void att_scores_kernel(const float* __restrict__ enc, const float* __restrict__ s1,
                                  const float* __restrict__ s0, const float* __restrict__ wavg,
                                  const float* __restrict__ bavg, const float* __restrict__ v,
                                  float* __restrict__ score, int B, long L, long La, int N, int ksz) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long o = warp; o < (long)B * La; o += nwarps) {
        const long b = o / La, la = o % La;
        float acc = 0.f;
        for (int c = lane; c < N; c += 32) {
            float avg = __ldg(bavg + c);
            for (int j = 0; j < ksz; ++j) {
                const float xn = enc[(b * L + la * ksz + j) * N + c] * s1[b * N + c] + s0[b * N + c];
                avg = fmaf(__ldg(wavg + c * ksz + j), xn, avg);
            }
            acc = fmaf(avg, v[b * N + c], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) score[o] = acc;
    }
}

// in-place softmax over each row of [B, La]; one CTA per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, long La) {
    __shared__ float red[32];
    __shared__ double scratch[32];
    float* row = x + (long)blockIdx.x * La;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float m = -INFINITY;
    for (long i = threadIdx.x; i < La; i += blockDim.x) m = fmaxf(m, row[i]);
    m = warp_max(m);
    if (lane == 0) red[wid] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    double s = 0.0;
    for (long i = threadIdx.x; i < La; i += blockDim.x) s += (double)expf(row[i] - m);
    s = block_sum(s, scratch);
    const float inv = (float)(1.0 / s);
    for (long i = threadIdx.x; i < La; i += blockDim.x) row[i] = expf(row[i] - m) * inv;
}

// rowscale[b,l] = 1 + sm[b, min(floor(l*scale), La-1)]   (ATen nearest-upsample index rule)
__global__ void att_rowscale_kernel(const float* __restrict__ sm, float* __restrict__ rowscale, int B, long L,
                                    long La, float scale) {
    const long total = (long)B * L;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / L, l = idx % L;
        long src;
        if (La == L) src = l;
        else if (L == 2 * La) src = l >> 1;
        else {
            src = (long)floorf(__fmul_rn((float)l, scale));
            if (src > La - 1) src = La - 1;
        }
        rowscale[idx] = 1.0f + sm[b * La + src];
    }
}

// ------------------------------------------------------------------------------------------
// speaker ResNet pieces
// ------------------------------------------------------------------------------------------
// per-channel partial sums over rows: partial[p][c] = {sum, sumsq}; blockDim = 256, C in {64,128,256}.  Thread = 4
// consecutive channels (float4 loads), four independent loads in flight, fp32 over 16 rows then fp64; kBnParts blocks
// (4 per SM: the 128-block scalar form of round 1 reached 0.4 TB/s).
constexpr int kBnParts = 592;
__global__ void __launch_bounds__(256) channel_stats_kernel(const float* __restrict__ y, double* __restrict__ partial,
                                                            long rows, int C) {
    __shared__ double sh[8][256];
    const int c4n = C / 4, lanes = 256 / c4n;
    const int c4 = threadIdx.x % c4n, rl = threadIdx.x / c4n;
    const long per = (rows + gridDim.x - 1) / gridDim.x;
    const long beg = blockIdx.x * per, end = min(beg + per, rows);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    double ds[4] = {0.0, 0.0, 0.0, 0.0}, dq[4] = {0.0, 0.0, 0.0, 0.0};
    long r = beg + rl;
    while (r < end) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
        auto acc = [&](const float4 v) {
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
        };
        int u = 0;
        for (; u < 4 && r + 3L * lanes < end; ++u, r += 4L * lanes) {
            const float4 v0 = ld_stream(y4 + r * c4n + c4), v1 = ld_stream(y4 + (r + lanes) * c4n + c4);
            const float4 v2 = ld_stream(y4 + (r + 2L * lanes) * c4n + c4), v3 = ld_stream(y4 + (r + 3L * lanes) * c4n + c4);
            acc(v0); acc(v1); acc(v2); acc(v3);
        }
        if (u < 4)
            for (; r < end; r += lanes) acc(ld_stream(y4 + r * c4n + c4));
        ds[0] += (double)s.x; ds[1] += (double)s.y; ds[2] += (double)s.z; ds[3] += (double)s.w;
        dq[0] += (double)q.x; dq[1] += (double)q.y; dq[2] += (double)q.z; dq[3] += (double)q.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { sh[k][threadIdx.x] = ds[k]; sh[4 + k][threadIdx.x] = dq[k]; }
    __syncthreads();
    if (rl == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double a = ds[k], b2 = dq[k];
            for (int i = 1; i < lanes; ++i) { a += sh[k][i * c4n + c4]; b2 += sh[4 + k][i * c4n + c4]; }
            partial[((long)blockIdx.x * C + c4 * 4 + k) * 2 + 0] = a;
            partial[((long)blockIdx.x * C + c4 * 4 + k) * 2 + 1] = b2;
        }
    }
}

// BatchNorm1d scale/shift. training: batch stats from partials (+ running update, momentum 0.1,
// unbiased variance); eval: running stats.
__global__ void bn_finalize_kernel(const double* __restrict__ partial, int nparts, double count,
                                   const float* __restrict__ weight, const float* __restrict__ bias,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, int C, int training,
                                   float eps, float momentum) {
    // one warp per channel: lane l adds every 32nd partial, then a shuffle tree (fixed order: deterministic)
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    float mean, var;
    if (training) {
        double s = 0.0, q = 0.0;
        for (int p = lane; p < nparts; p += 32) {
            s += partial[((long)p * C + c) * 2 + 0];
            q += partial[((long)p * C + c) * 2 + 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane != 0) return;
        const double m = s / count;
        double v = q / count - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m; var = (float)v;
        if (running_mean) {
            const double unb = count > 1.0 ? v * count / (count - 1.0) : v;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
        }
    } else {
        if (lane != 0) return;
        mean = running_mean[c]; var = running_var[c];
    }
    const float sc = weight[c] / sqrtf(var + eps);
    scale[c] = sc;
    shift[c] = bias[c] - mean * sc;
}

// out = prelu(y*scale[c] + shift[c])
__global__ void affine_prelu_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                    const float* __restrict__ shift, const float* __restrict__ prelu_a,
                                    float* __restrict__ out, long total4, int c4n) {
    const float a = __ldg(prelu_a);
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
        float4 v = reinterpret_cast<const float4*>(y)[idx];
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        v.x = v.x >= 0.f ? v.x : a * v.x; v.y = v.y >= 0.f ? v.y : a * v.y;
        v.z = v.z >= 0.f ? v.z : a * v.z; v.w = v.w >= 0.f ? v.w : a * v.w;
        reinterpret_cast<float4*>(out)[idx] = v;
    }
}

// out[b,lp,c] = max_{i<3} prelu(y[b,3lp+i,c]*scale[c] + shift[c] + skip[b,3lp+i,c])
__global__ void affine_add_prelu_pool3_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                              const float* __restrict__ shift, const float* __restrict__ skip,
                                              const float* __restrict__ prelu_a, float* __restrict__ out, int B,
                                              long Lin, long Lout, int c4n) {
    const float a = __ldg(prelu_a);
    const long total = (long)B * Lout * c4n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        const long r = idx / c4n;
        const long lp = r % Lout, b = r / Lout;
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const long src = (b * Lin + lp * 3 + i) * c4n + c4;
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

// emb[b,c] = (sum_l x[b,l,c]) / div[b]; one CTA (256 threads) per utterance, C in {64,128,256}
__global__ void __launch_bounds__(256) time_sum_kernel(const float* __restrict__ x, float* __restrict__ emb, long Lx,
                                                       int C, const float* __restrict__ div) {
    __shared__ double sh[256];
    const int lanes = 256 / C;
    const int c = threadIdx.x % C, rl = threadIdx.x / C;
    const float* xb = x + (long)blockIdx.x * Lx * C;
    double s = 0.0;
    for (long l = rl; l < Lx; l += lanes) s += (double)xb[l * C + c];
    sh[threadIdx.x] = s;
    __syncthreads();
    if (rl == 0) {
        for (int i = 1; i < lanes; ++i) s += sh[i * C + c];
        emb[(long)blockIdx.x * C + c] = (float)s / div[blockIdx.x];
    }
}

// out[b,n] (+)= bias[n] + sum_k in[b*ldin + k] * W[n*ldw + k]; one warp per output
__global__ void small_linear_kernel(const float* __restrict__ in, long ldin, const float* __restrict__ W, long ldw,
                                    const float* __restrict__ bias, float* __restrict__ out, long ldout, int B, int N,
                                    int K, int accumulate) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    if (warp >= (long)B * N) return;
    const long b = warp / N, n = warp % N;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(in[b * ldin + k], __ldg(W + n * ldw + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
        if (bias) acc += bias[n];
        if (accumulate) acc += out[b * ldout + n];
        out[b * ldout + n] = acc;
    }
}

static inline unsigned grid_for(long total, int threads) {
    long g = (total + threads - 1) / threads;
    const long cap = 148L * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dprnn

using namespace dprnn;

extern "C" {

int dprnn_encoder_fwd(const float* wave, const float* w, float* enc, int B, int T, int N, int ksz, int stride,
                      void* stream) {
    DPRNN_CHECK_ARG(wave && w && enc && B > 0 && N > 0 && N % 4 == 0 && ksz > 0 && stride > 0 && T >= ksz);
    const long L = (T - ksz) / stride + 1;
    const long total = (long)B * L * (N / 4);
    if (ksz == 2 && stride == 1 && 256 % (N / 4) == 0 && total < (1L << 31) && ((uintptr_t)w | (uintptr_t)enc) % 16 == 0) {
        encoder_k2s1_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(wave, w, (float4*)enc, (unsigned)B,
                                                                                    (unsigned)T, (unsigned)L, (unsigned)(N / 4));
        DPRNN_CHECK_LAUNCH();
        return 0;
    }
    encoder_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(wave, w, enc, B, T, L, N, ksz, stride);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

size_t dprnn_utt_stats_workspace_bytes(int B) { return (size_t)B * kStatParts * 2 * sizeof(double); }

int dprnn_utt_stats(const float* x, int B, long elems_per_utt, float eps, void* workspace, float* mean_rstd,
                    void* stream) {
    DPRNN_CHECK_ARG(x && workspace && mean_rstd && B > 0 && elems_per_utt > 0 && elems_per_utt % 4 == 0);
    DPRNN_CHECK_ARG(B <= 65535);
    dim3 grid(kStatParts, B);
    utt_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (double*)workspace, elems_per_utt / 4);
    DPRNN_CHECK_LAUNCH();
    utt_stats_finalize_kernel<<<B, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, mean_rstd, kStatParts,
                                                                 (double)elems_per_utt, (double)eps);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_norm_affine(const float* mean_rstd, const float* gamma, const float* beta, const float* mulc, float* s1,
                      float* s0, int B, int C, void* stream) {
    DPRNN_CHECK_ARG(mean_rstd && gamma && beta && s1 && s0 && B > 0 && C > 0);
    norm_affine_kernel<<<cdiv((long)B * C, 256), 256, 0, (cudaStream_t)stream>>>(mean_rstd, gamma, beta, mulc, s1, s0,
                                                                               B, C);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_prologue_apply(const float* a, float* out, long rows, int C, long rows_per_utt, const float* p_scale,
                         const float* p_shift, const float* p_add, const float* rowscale, void* stream) {
    DPRNN_CHECK_ARG(a && out && rows > 0 && C % 4 == 0 && rows_per_utt > 0);
    DPRNN_CHECK_ARG((p_scale == nullptr) == (p_shift == nullptr));
    prologue_apply_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        a, out, rows * (C / 4), C / 4, rows_per_utt, p_scale, p_shift, p_add, rowscale, nullptr);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_prologue_apply_ragged(const float* a, float* out, long rows, int C, const int* row_utt, const float* p_scale,
                                const float* p_shift, const float* p_add, const float* rowscale, void* stream) {
    DPRNN_CHECK_ARG(a && out && rows > 0 && C % 4 == 0 && row_utt);
    DPRNN_CHECK_ARG((p_scale == nullptr) == (p_shift == nullptr));
    prologue_apply_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        a, out, rows * (C / 4), C / 4, 1, p_scale, p_shift, p_add, rowscale, row_utt);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

#define DPRNN_CHECK_H16(h16) DPRNN_CHECK_ARG((h16) == DPRNN_H16_BF16 || (h16) == DPRNN_H16_FP16)

int dprnn_cast_h16(const float* x, void* out, long elems, int h16, void* stream) {
    DPRNN_CHECK_ARG(x && out && elems > 0 && elems % 4 == 0);
    DPRNN_CHECK_H16(h16);
    auto kern = h16 ? cast_bf16_kernel<true> : cast_bf16_kernel<false>;
    kern<<<grid_for(elems / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, (uint2*)out, elems / 4);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
int dprnn_cast_bf16(const float* x, void* out, long elems, void* stream) {
    return dprnn_cast_h16(x, out, elems, DPRNN_H16_BF16, stream);
}

int dprnn_norm_residual(const float* y, float* x, const float* mean_rstd, const float* gamma, const float* beta,
                        int B, long rows_per_utt, int C, void* x_bf16, void* stream) {
    DPRNN_CHECK_ARG(y && x && mean_rstd && gamma && beta && B > 0 && rows_per_utt > 0 && C % 4 == 0);
    const long per4 = rows_per_utt * (C / 4);
    norm_residual_kernel<false><<<grid_for(per4 * B, 256), 256, 0, (cudaStream_t)stream>>>(y, x, mean_rstd, gamma, beta,
                                                                                         per4 * B, per4, C / 4,
                                                                                         (uint2*)x_bf16);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_norm_residual_to(const float* y, const float* x, const float* mean_rstd, const float* gamma, const float* beta,
                           int B, long rows_per_utt, int C, float* x_out, void* x_bf16, void* stream) {
    DPRNN_CHECK_ARG(y && x && x_out && x_out != x && mean_rstd && gamma && beta && B > 0 && rows_per_utt > 0 && C % 4 == 0);
    const long per4 = rows_per_utt * (C / 4);
    norm_residual_kernel<false><<<grid_for(per4 * B, 256), 256, 0, (cudaStream_t)stream>>>(
        y, const_cast<float*>(x), mean_rstd, gamma, beta, per4 * B, per4, C / 4, (uint2*)x_bf16, x_out);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_norm_residual_yh16(const void* y_h16, float* x, const float* mean_rstd, const float* gamma, const float* beta,
                             int B, long rows_per_utt, int C, void* x_h16, int h16, void* stream) {
    DPRNN_CHECK_ARG(y_h16 && x && mean_rstd && gamma && beta && B > 0 && rows_per_utt > 0 && C % 8 == 0);
    DPRNN_CHECK_H16(h16);
    const long per8 = rows_per_utt * (C / 8);
    auto kern = h16 ? norm_residual_ybf16_kernel<true> : norm_residual_ybf16_kernel<false>;
    kern<<<grid_for(per8 * B, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)y_h16, x, mean_rstd, gamma, beta, per8 * B, per8, C / 8, (uint4*)x_h16);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
int dprnn_norm_residual_ybf16(const void* y_bf16, float* x, const float* mean_rstd, const float* gamma, const float* beta,
                              int B, long rows_per_utt, int C, void* x_bf16, void* stream) {
    return dprnn_norm_residual_yh16(y_bf16, x, mean_rstd, gamma, beta, B, rows_per_utt, C, x_bf16, DPRNN_H16_BF16, stream);
}

int dprnn_norm_residual_h16res(const void* y_h16, void* x_h16, float* x_f32_out, const float* mean_rstd,
                               const float* gamma, const float* beta, int B, long rows_per_utt, int C, int h16,
                               void* stream) {
    DPRNN_CHECK_ARG(y_h16 && x_h16 && mean_rstd && gamma && beta && B > 0 && rows_per_utt > 0 && C % 8 == 0);
    DPRNN_CHECK_H16(h16);
    const long per8 = rows_per_utt * (C / 8);
    auto kern = h16 ? norm_residual_bf16res_kernel<true> : norm_residual_bf16res_kernel<false>;
    kern<<<grid_for(per8 * B, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)y_h16, (uint4*)x_h16, x_f32_out, mean_rstd, gamma, beta, per8 * B, per8, C / 8);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
int dprnn_norm_residual_bf16res(const void* y_bf16, void* x_bf16, float* x_f32_out, const float* mean_rstd,
                                const float* gamma, const float* beta, int B, long rows_per_utt, int C, void* stream) {
    return dprnn_norm_residual_h16res(y_bf16, x_bf16, x_f32_out, mean_rstd, gamma, beta, B, rows_per_utt, C,
                                      DPRNN_H16_BF16, stream);
}

int dprnn_num_chunks(long L, int K, int P) { return (int)((L + K) / P + 1); }

int dprnn_unfold(const float* y, float* x, int B, long L, int K, int P, int F, void* stream) {
    DPRNN_CHECK_ARG(y && x && B > 0 && L > 0 && K > 0 && P > 0 && F % 4 == 0);
    const int S = dprnn_num_chunks(L, K, P);
    unfold_kernel<false><<<grid_for((long)B * S * K * (F / 4), 256), 256, 0, (cudaStream_t)stream>>>(y, x, B, L, S, K, P,
                                                                                                  F / 4, nullptr);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_unfold_h16(const float* y, float* x, void* x_h16, int B, long L, int K, int P, int F, int h16, void* stream) {
    DPRNN_CHECK_ARG(y && x_h16 && B > 0 && L > 0 && K > 0 && P > 0 && F % 4 == 0);      // x may be NULL
    DPRNN_CHECK_H16(h16);
    const int S = dprnn_num_chunks(L, K, P);
    const long rows = (long)B * S * K;
    if (!x && F % 8 == 0 && rows * (F / 8) < (1L << 31) && (long)S * P + K < (1L << 31) && L < (1L << 31) &&
        ((uintptr_t)y | (uintptr_t)x_h16) % 16 == 0) {
        auto fast = h16 ? unfold_h16_rows_kernel<true> : unfold_h16_rows_kernel<false>;
        fast<<<grid_for(rows * (F / 8), 256), 256, 0, (cudaStream_t)stream>>>(
            (const float4*)y, (uint4*)x_h16, (unsigned)rows, (unsigned)L, (unsigned)S, (unsigned)K, (unsigned)P,
            (unsigned)(F / 8));
        DPRNN_CHECK_LAUNCH();
        return 0;
    }
    auto kern = h16 ? unfold_kernel<true> : unfold_kernel<false>;
    kern<<<grid_for((long)B * S * K * (F / 4), 256), 256, 0, (cudaStream_t)stream>>>(y, x, B, L, S, K, P, F / 4,
                                                                                  (uint2*)x_h16);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
int dprnn_unfold_bf16(const float* y, float* x, void* x_bf16, int B, long L, int K, int P, int F, void* stream) {
    return dprnn_unfold_h16(y, x, x_bf16, B, L, K, P, F, DPRNN_H16_BF16, stream);
}

int dprnn_fold_prelu(const float* x, float* out, int B, long L, int K, int P, int F, const float* prelu_a,
                     void* stream) {
    DPRNN_CHECK_ARG(x && out && B > 0 && L > 0 && K > 0 && P > 0 && F % 4 == 0);
    const int S = dprnn_num_chunks(L, K, P);
    fold_prelu_kernel<<<grid_for((long)B * L * (F / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, out, B, L, S, K, P,
                                                                                           F / 4, prelu_a);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_norm_residual_fold_prelu_h16(const void* y_h16, const void* x_h16, const float* mean_rstd, const float* gamma,
                                       const float* beta, float* out, int B, long L, int K, int P, int F,
                                       const float* prelu_a, int h16, void* stream) {
    DPRNN_CHECK_ARG(y_h16 && x_h16 && mean_rstd && gamma && beta && out && B > 0 && L > 0 && K > 0 && P > 0 && F % 8 == 0);
    DPRNN_CHECK_ARG(((uintptr_t)y_h16 | (uintptr_t)x_h16 | (uintptr_t)out | (uintptr_t)gamma | (uintptr_t)beta) % 16 == 0);
    DPRNN_CHECK_H16(h16);
    const int S = dprnn_num_chunks(L, K, P);
    // 32-bit index arithmetic when the element counts of both spaces (+ one grid stride) fit
    const bool i32 = (long)B * L * (F / 8) < (1L << 31) && (long)B * S * K * (F / 8) < (1L << 31) && L + K < (1L << 31);
    auto kern = i32 ? (h16 ? norm_residual_fold_prelu_kernel<true, unsigned> : norm_residual_fold_prelu_kernel<false, unsigned>)
                    : (h16 ? norm_residual_fold_prelu_kernel<true, long> : norm_residual_fold_prelu_kernel<false, long>);
    kern<<<grid_for((long)B * L * (F / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)y_h16, (const uint4*)x_h16, mean_rstd, gamma, beta, out, B, L, S, K, P, F / 8, prelu_a);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_mask_decode(const float* mask, long mask_utt_stride, const float* enc, const float* wdec, float* out,
                      long out_utt_stride, int B, long L, int N, int ksz, int stride, void* stream) {
    DPRNN_CHECK_ARG(mask && enc && wdec && out && B > 0 && L > 0 && N > 0 && ksz > 0 && stride > 0);
    const int T = (int)((L - 1) * stride + ksz);
    if (ksz == 2 && stride == 1 && N == 64 && mask_utt_stride % 4 == 0 &&
        ((uintptr_t)mask | (uintptr_t)enc) % 16 == 0) {
        const int bpu = (int)cdiv(L + 1, 256);
        mask_decode_k2s1_kernel<<<(unsigned)((long)B * bpu), 256, 0, (cudaStream_t)stream>>>(
            mask, mask_utt_stride, enc, wdec, out, out_utt_stride, L, bpu);
        DPRNN_CHECK_LAUNCH();
        return 0;
    }
    mask_decode_kernel<<<grid_for((long)B * T * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        mask, mask_utt_stride, enc, wdec, out, out_utt_stride, B, L, T, N, ksz, stride);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_mask_apply(const float* mask, const float* enc, float* out, long elems, void* stream) {
    DPRNN_CHECK_ARG(mask && enc && out && elems > 0 && elems % 4 == 0);
    mask_apply_kernel<<<grid_for(elems / 4, 256), 256, 0, (cudaStream_t)stream>>>(mask, enc, out, elems / 4);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_att_rowscale(const float* enc, const float* s1, const float* s0, const float* wavg, const float* bavg,
                       const float* v, float* scores, float* rowscale, int B, long L, int N, int ksz, void* stream) {
    DPRNN_CHECK_ARG(enc && s1 && s0 && wavg && bavg && v && scores && rowscale && B > 0 && N > 0 && ksz > 0 && L >= ksz);
    const long La = (L - ksz) / ksz + 1;
    att_scores_kernel<<<grid_for((long)B * La * 32, 256), 256, 0, (cudaStream_t)stream>>>(enc, s1, s0, wavg, bavg, v,
                                                                                        scores, B, L, La, N, ksz);
    DPRNN_CHECK_LAUNCH();
    softmax_rows_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(scores, La);
    DPRNN_CHECK_LAUNCH();
    const float scale = (float)La / (float)L;   // ATen: static_cast<float>(input_size) / output_size
    att_rowscale_kernel<<<grid_for((long)B * L, 256), 256, 0, (cudaStream_t)stream>>>(scores, rowscale, B, L, La,
                                                                                    scale);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

size_t dprnn_bn_workspace_bytes(int C) { return (size_t)kBnParts * C * 2 * sizeof(double); }

int dprnn_batchnorm_affine(const float* y, long rows, int C, const float* weight, const float* bias,
                           float* running_mean, float* running_var, int training, float eps, float momentum,
                           void* workspace, float* scale, float* shift, void* stream) {
    DPRNN_CHECK_ARG(weight && bias && scale && shift && C > 0 && 256 % C == 0);
    if (training) {
        DPRNN_CHECK_ARG(y && workspace && rows > 0);
        DPRNN_CHECK_ARG(C % 4 == 0 && (uintptr_t)y % 16 == 0);
        channel_stats_kernel<<<kBnParts, 256, 0, (cudaStream_t)stream>>>(y, (double*)workspace, rows, C);
        DPRNN_CHECK_LAUNCH();
    } else {
        DPRNN_CHECK_ARG(running_mean && running_var);
    }
    bn_finalize_kernel<<<cdiv(C, 4), 128, 0, (cudaStream_t)stream>>>((const double*)workspace, kBnParts,
                                                                     (double)rows, weight, bias, running_mean,
                                                                     running_var, scale, shift, C, training, eps,
                                                                     momentum);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_affine_prelu(const float* y, const float* scale, const float* shift, const float* prelu_a, float* out,
                       long rows, int C, void* stream) {
    DPRNN_CHECK_ARG(y && scale && shift && prelu_a && out && rows > 0 && C % 4 == 0);
    affine_prelu_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(y, scale, shift, prelu_a, out,
                                                                                       rows * (C / 4), C / 4);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_affine_add_prelu_pool3(const float* y, const float* scale, const float* shift, const float* skip,
                                 const float* prelu_a, float* out, int B, long Lin, int C, void* stream) {
    DPRNN_CHECK_ARG(y && scale && shift && skip && prelu_a && out && B > 0 && Lin >= 3 && C % 4 == 0);
    const long Lout = Lin / 3;
    affine_add_prelu_pool3_kernel<<<grid_for((long)B * Lout * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        y, scale, shift, skip, prelu_a, out, B, Lin, Lout, C / 4);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_time_sum(const float* x, float* emb, int B, long Lx, int C, const float* div, void* stream) {
    DPRNN_CHECK_ARG(x && emb && div && B > 0 && Lx > 0 && C > 0 && 256 % C == 0);
    time_sum_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, emb, Lx, C, div);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_small_linear(const float* in, long ldin, const float* W, long ldw, const float* bias, float* out,
                       long ldout, int B, int N, int K, int accumulate, void* stream) {
    DPRNN_CHECK_ARG(in && W && out && B > 0 && N > 0 && K > 0);
    small_linear_kernel<<<cdiv((long)B * N * 32, 256), 256, 0, (cudaStream_t)stream>>>(in, ldin, W, ldw, bias, out,
                                                                                     ldout, B, N, K, accumulate);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
