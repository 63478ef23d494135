// Backward of the separation path (cfg 5: DPRNN-Spe training step), exact fp32 on CUDA cores - correctness first.
// Generic building blocks: weight-gradient contraction C = A^T B over millions of rows (two-stage deterministic
// reduction), column sums, GroupNorm(1,C) / gLN backward per utterance, BatchNorm1d (train) backward per channel, and the
// pointwise adjoints (PReLU, MaxPool1d(3), gated head, mask + decoder, encoder).  The data-gradient contractions
// dX = dY W reuse dprnn_gemm_f32 (the nn.Linear / Conv1d weight [N_out, K_in] is already the [K, N] operand it wants),
// fold / unfold are each other's adjoints (dprnn_unfold / dprnn_fold_prelu with a NULL slope).
#include "common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {

static inline unsigned bgrid(long total, int threads) {
    long g = (total + threads - 1) / threads;
    const long cap = 148L * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---------------------------------------------------------------- C[N1,N2] (+)= A[M,N1]^T B[M,N2]
constexpr int AT_TILE = 64, AT_RK = 16;
__global__ void __launch_bounds__(256) gemm_atb_partial_kernel(const float* __restrict__ A, long lda,
                                                               const float* __restrict__ B, long ldb, long M, int N1,
                                                               int N2, long rows_per_chunk, float* __restrict__ partial,
                                                               int vec) {
    __shared__ __align__(16) float As[AT_RK][AT_TILE], Bs[AT_RK][AT_TILE];
    const int i0 = blockIdx.x * AT_TILE, j0 = blockIdx.y * AT_TILE;
    const long r0 = (long)blockIdx.z * rows_per_chunk, r1 = min(r0 + rows_per_chunk, M);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;        // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;  // loader: row lr (0..15), 4 columns at lc
    for (long r = r0; r < r1; r += AT_RK) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (r + lr < r1) {
            if (vec) {
                if (i0 + lc < N1) av = *reinterpret_cast<const float4*>(A + (r + lr) * lda + i0 + lc);
                if (j0 + lc < N2) bv = *reinterpret_cast<const float4*>(B + (r + lr) * ldb + j0 + lc);
            } else {        // widths / strides that are not multiples of 4 (e.g. the 251 speaker logits): scalar loads
                const float* ap = A + (r + lr) * lda + i0 + lc;
                const float* bp = B + (r + lr) * ldb + j0 + lc;
                if (i0 + lc + 0 < N1) av.x = ap[0];
                if (i0 + lc + 1 < N1) av.y = ap[1];
                if (i0 + lc + 2 < N1) av.z = ap[2];
                if (i0 + lc + 3 < N1) av.w = ap[3];
                if (j0 + lc + 0 < N2) bv.x = bp[0];
                if (j0 + lc + 1 < N2) bv.y = bp[1];
                if (j0 + lc + 2 < N2) bv.z = bp[2];
                if (j0 + lc + 3 < N2) bv.w = bp[3];
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&As[lr][lc]) = av;
        *reinterpret_cast<float4*>(&Bs[lr][lc]) = bv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < AT_RK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(aa[a], bb[b], acc[a][b]);
        }
    }
    float* p = partial + (long)blockIdx.z * N1 * N2;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        if (i >= N1) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + tx * 4 + b;
            if (j < N2) p[(long)i * N2 + j] = acc[a][b];
        }
    }
}

__global__ void chunk_reduce_kernel(const float* __restrict__ partial, int chunks, long n, float* __restrict__ out,
                                    long ldo, int ncols, int accumulate) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < chunks; ++c) s += (double)partial[(long)c * n + idx];
        float* o = out + (idx / ncols) * ldo + (idx % ncols);
        *o = (accumulate ? *o : 0.f) + (float)s;
    }
}

// ---------------------------------------------------------------- out[n] (+)= sum_m X[m, n] (optionally * Y[m, n])
__global__ void __launch_bounds__(256) col_sum_partial_kernel(const float* __restrict__ X, long ldx,
                                                              const float* __restrict__ Y, long ldy, long M, int N,
                                                              float* __restrict__ partial) {
    // blockDim = 256 = 32 columns x 8 row lanes; grid (ceil(N/32), parts)
    __shared__ float sh[8][32];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5, c = blockIdx.x * 32 + cl;
    const long per = (M + gridDim.y - 1) / gridDim.y;
    const long r0 = (long)blockIdx.y * per, r1 = min(r0 + per, M);
    float s = 0.f;
    if (c < N)
        for (long r = r0 + rl; r < r1; r += 8) s += Y ? X[r * ldx + c] * Y[r * ldy + c] : X[r * ldx + c];
    sh[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && c < N) {
        for (int i = 1; i < 8; ++i) s += sh[i][cl];
        partial[(long)blockIdx.y * N + c] = s;
    }
}

// ---------------------------------------------------------------- GroupNorm(1,C) / gLN backward, per utterance
// z = gamma * yhat + beta, yhat = (y - mean_b) * rstd_b.  Given dz: dgamma_c += sum dz*yhat, dbeta_c += sum dz,
// dy = rstd_b * (dz*gamma - mean_b(dz*gamma) - yhat * mean_b(dz*gamma*yhat)).
// Kernel 1: per (utterance, part) partial {s1, s2} in fp64 and per-channel partial {dgamma, dbeta}.
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                                            const float* __restrict__ mean_rstd,
                                                            const float* __restrict__ gamma, long rows_per_utt, int C,
                                                            double* __restrict__ utt_part, float* __restrict__ ch_part) {
    // thread = 4 consecutive channels (float4 loads), 256 / (C/4) rows per pass; fp32 partials per thread (a few hundred
    // terms), fp64 across the block
    __shared__ double scratch[32];
    __shared__ float4 shg[256], shb[256];
    const int b = blockIdx.y, p = blockIdx.x, nparts = gridDim.x;
    const int c4n = C / 4, lanes = 256 / c4n, c4 = threadIdx.x % c4n, rl = threadIdx.x / c4n;
    const float mean = mean_rstd[2 * b], rstd = mean_rstd[2 * b + 1];
    const float4 g = reinterpret_cast<const float4*>(gamma)[c4];
    const long per = (rows_per_utt + nparts - 1) / nparts;
    const long r0 = (long)p * per, r1 = min(r0 + per, rows_per_utt);
    const float4* dzb = reinterpret_cast<const float4*>(dz + (long)b * rows_per_utt * C);
    const float4* yb = reinterpret_cast<const float4*>(y + (long)b * rows_per_utt * C);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    float f1 = 0.f, f2 = 0.f;
    if (rl < lanes) {
        for (long r = r0 + rl; r < r1; r += lanes) {
            const float4 d = ld_stream(dzb + r * c4n + c4), yv = ld_stream(yb + r * c4n + c4);
            const float4 yh = make_float4((yv.x - mean) * rstd, (yv.y - mean) * rstd, (yv.z - mean) * rstd, (yv.w - mean) * rstd);
            dg.x = fmaf(d.x, yh.x, dg.x); dg.y = fmaf(d.y, yh.y, dg.y); dg.z = fmaf(d.z, yh.z, dg.z); dg.w = fmaf(d.w, yh.w, dg.w);
            db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
            const float4 e = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
            f1 += (e.x + e.y) + (e.z + e.w);
            f2 += (e.x * yh.x + e.y * yh.y) + (e.z * yh.z + e.w * yh.w);
        }
    }
    shg[threadIdx.x] = dg; shb[threadIdx.x] = db;
    double s1 = block_sum((double)f1, scratch);
    double s2 = block_sum((double)f2, scratch);
    __syncthreads();
    if (threadIdx.x == 0) {
        utt_part[((long)b * nparts + p) * 2 + 0] = s1;
        utt_part[((long)b * nparts + p) * 2 + 1] = s2;
    }
    if (rl == 0) {
        for (int i = 1; i < lanes; ++i) {
            const float4 a = shg[i * c4n + c4], c = shb[i * c4n + c4];
            dg.x += a.x; dg.y += a.y; dg.z += a.z; dg.w += a.w;
            db.x += c.x; db.y += c.y; db.z += c.z; db.w += c.w;
        }
        float* o = ch_part + (((long)b * nparts + p) * C + c4 * 4) * 2;
        o[0] = dg.x; o[1] = db.x; o[2] = dg.y; o[3] = db.y; o[4] = dg.z; o[5] = db.z; o[6] = dg.w; o[7] = db.w;
    }
}

// Kernel 2: one warp per channel (blocks 0 .. C/4-1) and one warp per utterance (the blocks after them) add the partials
// in a fixed order - lane l takes every 32nd partial, then a shuffle tree - so the result does not depend on scheduling.
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__global__ void __launch_bounds__(128) gn_bwd_finalize_kernel(const double* __restrict__ utt_part,
                                                              const float* __restrict__ ch_part, int B, int nparts, int C,
                                                              double count, float* __restrict__ s12,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cblocks = (C + 3) / 4;
    if ((int)blockIdx.x < cblocks) {
        const int c = blockIdx.x * 4 + warp;
        if (c >= C) return;
        double dg = 0.0, db = 0.0;
        for (long bp = lane; bp < (long)B * nparts; bp += 32) {
            const float2 v = *reinterpret_cast<const float2*>(ch_part + (bp * C + c) * 2);
            dg += v.x; db += v.y;
        }
        dg = warp_sum_f64(dg); db = warp_sum_f64(db);
        if (lane == 0) { dgamma[c] += (float)dg; dbeta[c] += (float)db; }
    } else {
        const int b = ((int)blockIdx.x - cblocks) * 4 + warp;
        if (b >= B) return;
        double s1 = 0.0, s2 = 0.0;
        for (int p = lane; p < nparts; p += 32) {
            s1 += utt_part[((long)b * nparts + p) * 2];
            s2 += utt_part[((long)b * nparts + p) * 2 + 1];
        }
        s1 = warp_sum_f64(s1); s2 = warp_sum_f64(s2);
        if (lane == 0) { s12[2 * b] = (float)(s1 / count); s12[2 * b + 1] = (float)(s2 / count); }
    }
}

__global__ void gn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                    const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                    const float* __restrict__ s12, float* __restrict__ dy, long total4, long per_utt4,
                                    int c4n, int accumulate, uint2* __restrict__ dy_bf16 = nullptr) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / per_utt4;
        const int c4 = (int)(idx % c4n);
        const float mean = mean_rstd[2 * b], rstd = mean_rstd[2 * b + 1], m1 = s12[2 * b], m2 = s12[2 * b + 1];
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 d = reinterpret_cast<const float4*>(dz)[idx];
        const float4 yv = reinterpret_cast<const float4*>(y)[idx];
        float4 o;
        o.x = rstd * (d.x * g.x - m1 - (yv.x - mean) * rstd * m2);
        o.y = rstd * (d.y * g.y - m1 - (yv.y - mean) * rstd * m2);
        o.z = rstd * (d.z * g.z - m1 - (yv.z - mean) * rstd * m2);
        o.w = rstd * (d.w * g.w - m1 - (yv.w - mean) * rstd * m2);
        if (accumulate) {
            const float4 a = reinterpret_cast<const float4*>(dy)[idx];
            o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        if (dy) reinterpret_cast<float4*>(dy)[idx] = o;          // dy == NULL: only the bf16 copy is wanted
        if (dy_bf16) {          // bf16 copy: the operand of the Linear's weight-gradient pass (dprnn_gemm_atb_dual, bf16 form)
            const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            dy_bf16[idx] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
    }
}

// ---------------------------------------------------------------- pointwise adjoints
// y = prelu(x): dx = dy * (x > 0 ? 1 : a) (optionally += ), partial da = sum dy * x * [x <= 0]
__global__ void __launch_bounds__(256) prelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                        const float* __restrict__ a_ptr, float* __restrict__ dx,
                                                        long total, double* __restrict__ da_part) {
    __shared__ double scratch[32];
    const float a = a_ptr[0];
    double s = 0.0;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const float xv = x[idx], d = dy[idx];
        dx[idx] = xv > 0.f ? d : a * d;
        if (!(xv > 0.f)) s += (double)(d * xv);
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) da_part[blockIdx.x] = s;
}

__global__ void sum_parts_kernel(const double* __restrict__ part, int n, float* __restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += part[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] += (float)s;
}

// y[b,lp,c] = max_i v[b,3lp+i,c]: dv[argmax] = dy (first maximum, as ATen), 0 elsewhere; rows beyond 3*Lout get 0
__global__ void pool3_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ v, float* __restrict__ dv, int B,
                                 long Lin, long Lout, int C) {
    const long total = (long)B * Lin * C;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const long r = idx / C, l = r % Lin, b = r / Lin;
        const long lp = l / 3;
        float g = 0.f;
        if (lp < Lout) {
            const float* vb = v + ((b * Lin + lp * 3) * C + c);
            const float v0 = vb[0], v1 = vb[C], v2 = vb[2 * C];
            int am = 0; float m = v0;
            if (v1 > m) { m = v1; am = 1; }
            if (v2 > m) { m = v2; am = 2; }
            if (am == (int)(l - lp * 3)) g = dy[(b * Lout + lp) * C + c];
        }
        dv[idx] = g;
    }
}

// out = tanh(po) * sigmoid(pg) with pre = [po | pg] ([rows, 2F], the packed 'og' layout is NOT used here: plain halves)
// dpre = [dg * sig(pg) * (1 - tanh(po)^2) | dg * tanh(po) * sig(pg) * (1 - sig(pg))]
__global__ void gated_bwd_kernel(const float* __restrict__ dgv, const float* __restrict__ pre, float* __restrict__ dpre,
                                 long rows, int F) {
    const long total = rows * F;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / F; const int c = (int)(idx % F);
        const float po = pre[r * 2 * F + c], pg = pre[r * 2 * F + F + c], d = dgv[idx];
        const float th = tanhf(po), sg = sigmoid_acc(pg);
        dpre[r * 2 * F + c] = d * sg * (1.f - th * th);
        dpre[r * 2 * F + F + c] = d * th * sg * (1.f - sg);
    }
}

__global__ void gated_fwd_kernel(const float* __restrict__ pre, float* __restrict__ out, long rows, int F) {
    const long total = rows * F;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / F; const int c = (int)(idx % F);
        out[idx] = tanhf(pre[r * 2 * F + c]) * sigmoid_acc(pre[r * 2 * F + F + c]);
    }
}

// elementwise helpers: out = a * b ; out (+)= alpha * a ; dpre = dy * act'(y)
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = a[i] * b[i];
}
__global__ void axpy_kernel(const float* __restrict__ a, float alpha, float* __restrict__ out, long n, int accumulate) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = (accumulate ? out[i] : 0.f) + alpha * a[i];
}
// act: 1 = relu (y > 0), 2 = sigmoid (y (1 - y)) with y the activation OUTPUT
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dpre, long n, int act) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float yv = y[i];
        dpre[i] = dy[i] * (act == 1 ? (yv > 0.f ? 1.f : 0.f) : yv * (1.f - yv));
    }
}

// decoder (ConvTranspose1d N->1, k=2, s=1) adjoint: dz[b,l,c] = dest[b,l] w[c,0] + dest[b,l+1] w[c,1]
__global__ void decoder_bwd_kernel(const float* __restrict__ dest, const float* __restrict__ w, float* __restrict__ dz,
                                   int B, long L, int N) {
    const long total = (long)B * L * N;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % N);
        const long r = idx / N, l = r % L, b = r / L;
        const float* d = dest + b * (L + 1) + l;
        dz[idx] = d[0] * __ldg(w + 2 * c) + d[1] * __ldg(w + 2 * c + 1);
    }
}
// dw_dec[c,j] = sum_{b,l} z[b,l,c] dest[b,l+j]: partial per block row-range; z = mask*enc
__global__ void __launch_bounds__(256) convw2_partial_kernel(const float* __restrict__ z, const float* __restrict__ sig,
                                                             int B, long L, long T, int N, float* __restrict__ partial) {
    // thread = (channel c, row lane); partial[blockIdx.x][c][2]
    __shared__ float sh[2][256];
    const int lanes = 256 / N, c = threadIdx.x % N, rl = threadIdx.x / N;
    const long rows = (long)B * L, per = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * per, r1 = min(r0 + per, rows);
    float a0 = 0.f, a1 = 0.f;
    long r = r0 + rl;
    for (; r + 3L * lanes < r1; r += 4L * lanes) {          // four independent loads in flight
        float zv[4], s0[4], s1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long rr = r + (long)u * lanes, b = rr / L, l = rr % L;
            zv[u] = z[rr * N + c]; s0[u] = __ldg(sig + b * T + l); s1[u] = __ldg(sig + b * T + l + 1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { a0 = fmaf(zv[u], s0[u], a0); a1 = fmaf(zv[u], s1[u], a1); }
    }
    for (; r < r1; r += lanes) {
        const long b = r / L, l = r % L;
        const float zv = z[r * N + c];
        a0 = fmaf(zv, sig[b * T + l], a0);
        a1 = fmaf(zv, sig[b * T + l + 1], a1);
    }
    sh[0][threadIdx.x] = a0; sh[1][threadIdx.x] = a1;
    __syncthreads();
    if (rl == 0) {
        for (int i = 1; i < lanes; ++i) { a0 += sh[0][i * N + c]; a1 += sh[1][i * N + c]; }
        partial[((long)blockIdx.x * N + c) * 2] = a0;
        partial[((long)blockIdx.x * N + c) * 2 + 1] = a1;
    }
}
// encoder adjoint wrt the waveform is not needed (inputs carry no gradient); its weight gradient is convw2 with
// z := denc * [enc > 0] and sig := the waveform.

// per-utterance, per-channel sum over time: out[b,c] = sum_l X[b,l,c] (* Y[b,l,c]) - the FiLM / time-mean adjoints.
// grid (UCS_PARTS, B): partial sums in fp64, then a fixed-order finalize.
constexpr int UCS_PARTS = 32;
__global__ void __launch_bounds__(256) utt_col_sum_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                          long L, int C, double* __restrict__ partial) {
    __shared__ double sh[256];
    const int lanes = 256 / C, c = threadIdx.x % C, rl = threadIdx.x / C;
    const float* xb = X + (long)blockIdx.y * L * C;
    const float* yb = Y ? Y + (long)blockIdx.y * L * C : nullptr;
    const long per = (L + UCS_PARTS - 1) / UCS_PARTS;
    const long l0 = blockIdx.x * per, l1 = min(l0 + per, L);
    double s = 0.0;
    for (long l = l0 + rl; l < l1; l += lanes) s += (double)(yb ? xb[l * C + c] * yb[l * C + c] : xb[l * C + c]);
    sh[threadIdx.x] = s;
    __syncthreads();
    if (rl == 0) {
        for (int i = 1; i < lanes; ++i) s += sh[i * C + c];
        partial[((long)blockIdx.y * UCS_PARTS + blockIdx.x) * C + c] = s;
    }
}
__global__ void utt_col_sum_final_kernel(const double* __restrict__ partial, int C, long total, float* __restrict__ out) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long b = e / C;
    const int c = (int)(e % C);
    double s = 0.0;
    for (int p = 0; p < UCS_PARTS; ++p) s += partial[(b * UCS_PARTS + p) * C + c];
    out[e] = (float)s;
}

// ---------------------------------------------------------------- attention fusion backward (dprnn_spe.py:177-183,217-225)
// fused[b,l,c] = n[b,l,c] * v[b,c] * r[b,l],  r = 1 + softmax_j(s)[src(l)],  s[b,j] = sum_c v[b,c] * (bavg[c] + sum_i
// wavg[c,i] n[b, j*ksz+i, c]).   dr[row] = sum_c dfused * n * v                                   (warp per row)
__global__ void row_dot3_kernel(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ v,
                                long rows, long rows_per_utt, int C, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r = warp; r < rows; r += nwarps) {
        const float* vb = v + (r / rows_per_utt) * C;
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) acc = fmaf(A[r * C + c] * Bm[r * C + c], vb[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) out[r] = acc;
    }
}

__device__ __forceinline__ long att_src(long l, long L, long La, float scale) {      // ATen nearest-upsample index rule
    if (La == L) return l;
    if (L == 2 * La) return l >> 1;
    long src = (long)floorf(__fmul_rn((float)l, scale));
    return src > La - 1 ? La - 1 : src;
}

// one CTA per utterance: da[j] = sum_{l : src(l) = j} dr[l] (upsample adjoint, fixed order);  ds = a * (da - <a, da>)
// (softmax adjoint);  w2[l] = ds[l / ksz] for l < La*ksz, else 0 (what reaches the frames through the average conv)
__global__ void __launch_bounds__(256) att_softmax_bwd_kernel(const float* __restrict__ dr, const float* __restrict__ a,
                                                              long L, long La, int ksz, float scale,
                                                              float* __restrict__ w2, float* __restrict__ ds) {
    __shared__ double scratch[32];
    const long b = blockIdx.x;
    const float* drb = dr + b * L;
    const float* ab = a + b * La;
    float* dsb = ds + b * La;
    float* w2b = w2 + b * L;
    const float inv = (float)L / (float)La;
    double dot = 0.0;
    for (long j = threadIdx.x; j < La; j += 256) {
        long lo = (long)floorf((float)j * inv) - 2, hi = (long)ceilf((float)(j + 1) * inv) + 2;
        if (lo < 0) lo = 0;
        if (hi > L - 1 || j == La - 1) hi = L - 1;
        float s = 0.f;
        for (long l = lo; l <= hi; ++l)
            if (att_src(l, L, La, scale) == j) s += drb[l];
        dsb[j] = s;
        dot += (double)ab[j] * (double)s;
    }
    dot = block_sum(dot, scratch);
    const float fdot = (float)dot;
    for (long j = threadIdx.x; j < La; j += 256) {
        const float v = ab[j] * (dsb[j] - fdot);
        dsb[j] = v;
        for (int i = 0; i < ksz; ++i) w2b[j * ksz + i] = v;
    }
    for (long l = La * ksz + threadIdx.x; l < L; l += 256) w2b[l] = 0.f;
}

// g = dfused * r[row] + w2[row] * wavg[c, l % ksz];   dn = v[b,c] * g;   tdv = n * g  (summed over time by the caller)
__global__ void att_bwd_apply_kernel(const float* __restrict__ dfused, const float* __restrict__ n,
                                     const float* __restrict__ v, const float* __restrict__ r,
                                     const float* __restrict__ w2, const float* __restrict__ wavg, long L, int C, int ksz,
                                     long total, float* __restrict__ dn, float* __restrict__ tdv) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const long row = idx / C, b = row / L, l = row - b * L;
        const float g = fmaf(w2[row], wavg[c * ksz + (int)(l % ksz)], dfused[idx] * r[row]);
        dn[idx] = v[b * C + c] * g;
        tdv[idx] = n[idx] * g;
    }
}

// out[b,l,c] (+)= v[b,c] * (X ? X[b,l,c] : 1)   (broadcast of a per-utterance vector over time)
__global__ void bcast_mul_kernel(const float* __restrict__ v, const float* __restrict__ X, float* __restrict__ out, long L,
                                 int C, long total, int accumulate) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const long b = idx / (L * C);
        const float val = v[b * C + c] * (X ? X[idx] : 1.f);
        out[idx] = (accumulate ? out[idx] : 0.f) + val;
    }
}

// BatchNorm1d (train) backward apply: dy = scale_c * (dout - m1_c - yhat * m2_c), yhat = (y - mean_c) * rstd_c
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ m1,
                                    const float* __restrict__ m2, float* __restrict__ dy, long total, int C) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const float yh = (y[idx] - mean[c]) * rstd[c];
        dy[idx] = gamma[c] * rstd[c] * (dout[idx] - m1[c] - yh * m2[c]);
    }
}

// h_prev of the recurrence for the W_hh gradient: out[row(n,t), d*H + j] = h[row(n, t -/+ 1), d*H + j] (the step the
// forward of direction d visited before t), 0 at the first step.  Rows follow the forward's sequence geometry.
__global__ void shift_rows_kernel(const float* __restrict__ h, float* __restrict__ out, long nseq, int T, long seq_div,
                                  long seq_outer, long seq_inner, long step_stride, int Hd, int ndir) {
    const int h4n = ndir * Hd / 4, d4 = Hd / 4;
    const long total = nseq * T * h4n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % h4n);
        const long nt = idx / h4n;
        const int t = (int)(nt % T);
        const long n = nt / T;
        const int dir = c4 / d4;
        const long base = (n / seq_div) * seq_outer + (n % seq_div) * seq_inner;
        const int tp = dir ? t + 1 : t - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tp >= 0 && tp < T) v = reinterpret_cast<const float4*>(h)[(base + (long)tp * step_stride) * h4n + c4];
        reinterpret_cast<float4*>(out)[(base + (long)t * step_stride) * h4n + c4] = v;
    }
}

}  // namespace dprnn

using namespace dprnn;

extern "C" {

int dprnn_shift_rows(const float* h, float* out, long nseq, int T, long seq_div, long seq_outer_stride,
                     long seq_inner_stride, long step_stride, int hidden, int ndir, void* stream) {
    DPRNN_CHECK_ARG(h && out && nseq > 0 && T > 0 && seq_div > 0 && hidden % 4 == 0 && (ndir == 1 || ndir == 2));
    shift_rows_kernel<<<bgrid(nseq * T * (ndir * hidden / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        h, out, nseq, T, seq_div, seq_outer_stride, seq_inner_stride, step_stride, hidden, ndir);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

size_t dprnn_gemm_atb_workspace_bytes(long M, int N1, int N2) {
    const long chunks = (M + 8191) / 8192 < 1 ? 1 : ((M + 8191) / 8192 > 512 ? 512 : (M + 8191) / 8192);
    return (size_t)chunks * N1 * N2 * sizeof(float);
}

int dprnn_gemm_atb(const float* A, long lda, const float* B, long ldb, float* C, long ldc, long M, int N1, int N2,
                   int accumulate, void* workspace, void* stream) {
    DPRNN_CHECK_ARG(A && B && C && workspace && M > 0 && N1 > 0 && N2 > 0);
    const int vec = (N1 % 4 == 0 && N2 % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)A | (uintptr_t)B) % 16 == 0) ? 1 : 0;
    long chunks = (M + 8191) / 8192;
    chunks = chunks < 1 ? 1 : (chunks > 512 ? 512 : chunks);
    const long rpc = ((M + chunks - 1) / chunks + AT_RK - 1) / AT_RK * AT_RK;
    dim3 grid(cdiv(N1, AT_TILE), cdiv(N2, AT_TILE), (unsigned)chunks);
    gemm_atb_partial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, M, N1, N2, rpc, (float*)workspace, vec);
    DPRNN_CHECK_LAUNCH();
    chunk_reduce_kernel<<<bgrid((long)N1 * N2, 256), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, (int)chunks,
                                                                                    (long)N1 * N2, C, ldc, N2, accumulate);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

size_t dprnn_col_sum_workspace_bytes(int N) { return (size_t)128 * N * sizeof(float); }

int dprnn_col_sum(const float* X, long ldx, const float* Y, long ldy, long M, int N, float* out, int accumulate,
                  void* workspace, void* stream) {
    DPRNN_CHECK_ARG(X && out && workspace && M > 0 && N > 0);
    dim3 grid(cdiv(N, 32), 128);
    col_sum_partial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, Y, ldy, M, N, (float*)workspace);
    DPRNN_CHECK_LAUNCH();
    chunk_reduce_kernel<<<bgrid(N, 256), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, 128, N, out, N, N, accumulate);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

size_t dprnn_gn_bwd_workspace_bytes(int B, int C) { return (size_t)B * 64 * (2 * sizeof(double) + 2 * C * sizeof(float)) + (size_t)B * 2 * sizeof(float); }

static int gn_bwd_impl(const float* dz, const float* y, const float* mean_rstd, const float* gamma, int B,
                       long rows_per_utt, int C, float* dy, int accumulate_dy, float* dgamma, float* dbeta,
                       void* workspace, void* dy_bf16, void* stream) {
    DPRNN_CHECK_ARG(dz && y && mean_rstd && gamma && (dy || (dy_bf16 && !accumulate_dy)) && dgamma && dbeta && workspace && B > 0 &&
                    B <= 65535);
    DPRNN_CHECK_ARG(rows_per_utt > 0 && C % 4 == 0 && 256 % C == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int nparts = 64;
    double* utt_part = (double*)workspace;
    float* ch_part = (float*)(utt_part + (size_t)B * nparts * 2);
    float* s12 = ch_part + (size_t)B * nparts * C * 2;
    dim3 grid(nparts, B);
    gn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(dz, y, mean_rstd, gamma, rows_per_utt, C, utt_part, ch_part);
    DPRNN_CHECK_LAUNCH();
    gn_bwd_finalize_kernel<<<cdiv(C, 4) + cdiv(B, 4), 128, 0, st>>>(utt_part, ch_part, B, nparts, C,
                                                                   (double)rows_per_utt * C, s12, dgamma, dbeta);
    DPRNN_CHECK_LAUNCH();
    const long per4 = rows_per_utt * (C / 4);
    gn_bwd_apply_kernel<<<bgrid(per4 * B, 256), 256, 0, st>>>(dz, y, mean_rstd, gamma, s12, dy, per4 * B, per4, C / 4,
                                                             accumulate_dy, (uint2*)dy_bf16);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_groupnorm_bwd(const float* dz, const float* y, const float* mean_rstd, const float* gamma, int B,
                        long rows_per_utt, int C, float* dy, int accumulate_dy, float* dgamma, float* dbeta,
                        void* workspace, void* stream) {
    return gn_bwd_impl(dz, y, mean_rstd, gamma, B, rows_per_utt, C, dy, accumulate_dy, dgamma, dbeta, workspace, nullptr, stream);
}

int dprnn_groupnorm_bwd_h16(const float* dz, const float* y, const float* mean_rstd, const float* gamma, int B,
                            long rows_per_utt, int C, float* dy, int accumulate_dy, float* dgamma, float* dbeta,
                            void* workspace, void* dy_bf16, void* stream) {
    DPRNN_CHECK_ARG(dy_bf16 && (uintptr_t)dy_bf16 % 16 == 0);
    return gn_bwd_impl(dz, y, mean_rstd, gamma, B, rows_per_utt, C, dy, accumulate_dy, dgamma, dbeta, workspace, dy_bf16, stream);
}

int dprnn_prelu_bwd(const float* dy, const float* x, const float* prelu_a, float* dx, long n, float* da, void* workspace,
                    void* stream) {
    DPRNN_CHECK_ARG(dy && x && prelu_a && dx && da && workspace && n > 0);
    const unsigned grid = bgrid(n, 256);
    prelu_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, x, prelu_a, dx, n, (double*)workspace);
    DPRNN_CHECK_LAUNCH();
    sum_parts_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, (int)grid, da);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_pool3_bwd(const float* dy, const float* v, float* dv, int B, long Lin, int C, void* stream) {
    DPRNN_CHECK_ARG(dy && v && dv && B > 0 && Lin >= 3 && C > 0);
    pool3_bwd_kernel<<<bgrid((long)B * Lin * C, 256), 256, 0, (cudaStream_t)stream>>>(dy, v, dv, B, Lin, Lin / 3, C);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_gated_bwd(const float* dg, const float* pre, float* dpre, long rows, int F, void* stream) {
    DPRNN_CHECK_ARG(dg && pre && dpre && rows > 0 && F > 0);
    gated_bwd_kernel<<<bgrid(rows * F, 256), 256, 0, (cudaStream_t)stream>>>(dg, pre, dpre, rows, F);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_gated_fwd(const float* pre, float* out, long rows, int F, void* stream) {
    DPRNN_CHECK_ARG(pre && out && rows > 0 && F > 0);
    gated_fwd_kernel<<<bgrid(rows * F, 256), 256, 0, (cudaStream_t)stream>>>(pre, out, rows, F);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_mul(const float* a, const float* b, float* out, long n, void* stream) {
    DPRNN_CHECK_ARG(a && b && out && n > 0);
    mul_kernel<<<bgrid(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_axpy(const float* a, float alpha, float* out, long n, int accumulate, void* stream) {
    DPRNN_CHECK_ARG(a && out && n > 0);
    axpy_kernel<<<bgrid(n, 256), 256, 0, (cudaStream_t)stream>>>(a, alpha, out, n, accumulate);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_act_bwd(const float* dy, const float* y, float* dpre, long n, int act, void* stream) {
    DPRNN_CHECK_ARG(dy && y && dpre && n > 0 && (act == 1 || act == 2));
    act_bwd_kernel<<<bgrid(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, dpre, n, act);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_decoder_bwd(const float* dest, const float* wdec, float* dz, int B, long L, int N, void* stream) {
    DPRNN_CHECK_ARG(dest && wdec && dz && B > 0 && L > 0 && N > 0);
    decoder_bwd_kernel<<<bgrid((long)B * L * N, 256), 256, 0, (cudaStream_t)stream>>>(dest, wdec, dz, B, L, N);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

constexpr int kConvwParts = 592;      // 4 blocks per SM
size_t dprnn_convw2_workspace_bytes(int N) { return (size_t)kConvwParts * N * 2 * sizeof(float); }

/* dw[c, j] (+)= sum_{b,l} z[b,l,c] * sig[b, l + j], j in {0,1}: the weight gradient of the kernel-2 stride-1 decoder
 * (z = mask*enc, sig = d est) and encoder (z = d enc * [enc > 0], sig = waveform). */
int dprnn_convw2_grad(const float* z, const float* sig, int B, long L, int N, float* dw, int accumulate, void* workspace,
                      void* stream) {
    DPRNN_CHECK_ARG(z && sig && dw && workspace && B > 0 && L > 0 && N > 0 && 256 % N == 0);
    convw2_partial_kernel<<<kConvwParts, 256, 0, (cudaStream_t)stream>>>(z, sig, B, L, L + 1, N, (float*)workspace);
    DPRNN_CHECK_LAUNCH();
    chunk_reduce_kernel<<<bgrid(2L * N, 256), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, kConvwParts, 2L * N, dw, 2L * N,
                                                                              2 * N, accumulate);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

size_t dprnn_utt_col_sum_workspace_bytes(int B, int C) { return (size_t)B * UCS_PARTS * C * sizeof(double); }

int dprnn_utt_col_sum(const float* X, const float* Y, int B, long L, int C, float* out, void* workspace, void* stream) {
    DPRNN_CHECK_ARG(X && out && workspace && B > 0 && B <= 65535 && L > 0 && C > 0 && 256 % C == 0);
    utt_col_sum_kernel<<<dim3(UCS_PARTS, B), 256, 0, (cudaStream_t)stream>>>(X, Y, L, C, (double*)workspace);
    DPRNN_CHECK_LAUNCH();
    const long total = (long)B * C;
    utt_col_sum_final_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const double*)workspace, C,
                                                                                              total, out);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_row_dot3(const float* A, const float* Bm, const float* v, long rows, long rows_per_utt, int C, float* out,
                   void* stream) {
    DPRNN_CHECK_ARG(A && Bm && v && out && rows > 0 && rows_per_utt > 0 && rows % rows_per_utt == 0 && C > 0);
    row_dot3_kernel<<<bgrid(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(A, Bm, v, rows, rows_per_utt, C, out);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_att_softmax_bwd(const float* dr, const float* a, int B, long L, int ksz, float* w2, float* ds, void* stream) {
    DPRNN_CHECK_ARG(dr && a && w2 && ds && B > 0 && ksz > 0 && L >= ksz);
    const long La = (L - ksz) / ksz + 1;
    att_softmax_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(dr, a, L, La, ksz, (float)La / (float)L, w2, ds);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_att_bwd_apply(const float* dfused, const float* n, const float* v, const float* r, const float* w2,
                        const float* wavg, int B, long L, int C, int ksz, float* dn, float* tdv, void* stream) {
    DPRNN_CHECK_ARG(dfused && n && v && r && w2 && wavg && dn && tdv && B > 0 && L > 0 && C > 0 && ksz > 0);
    const long total = (long)B * L * C;
    att_bwd_apply_kernel<<<bgrid(total, 256), 256, 0, (cudaStream_t)stream>>>(dfused, n, v, r, w2, wavg, L, C, ksz, total,
                                                                            dn, tdv);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_bcast_mul(const float* v, const float* X, float* out, int B, long L, int C, int accumulate, void* stream) {
    DPRNN_CHECK_ARG(v && out && B > 0 && L > 0 && C > 0);
    bcast_mul_kernel<<<bgrid((long)B * L * C, 256), 256, 0, (cudaStream_t)stream>>>(v, X, out, L, C, (long)B * L * C, accumulate);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

int dprnn_bn_bwd_apply(const float* dout, const float* y, const float* mean, const float* rstd, const float* gamma,
                       const float* m1, const float* m2, float* dy, long rows, int C, void* stream) {
    DPRNN_CHECK_ARG(dout && y && mean && rstd && gamma && m1 && m2 && dy && rows > 0 && C > 0);
    bn_bwd_apply_kernel<<<bgrid(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(dout, y, mean, rstd, gamma, m1, m2, dy, rows * C, C);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
