// RawNet3 front-end of DPRNN-RawNet (src/models/rawnet/RawNet3.py:23-32,76-83; RawNetBasicBlock.py:8-28), cfg 4:
//   PreEmphasis (y[t] = x[t] - 0.97 x[t-1], reflect pad) -> InstanceNorm1d(1, eps 1e-4, affine)
//   -> parameterised sinc filterbank (256 filters x 251 taps, stride 10) -> |.| -> log(. + 1e-6) -> minus the time mean.
// Output is channels-last [B, T', 256] like every other activation of the path.
// The filterbank arithmetic restates asteroid_filterbanks==0.4.0 ParamSincFB (third-party, absent from the reference
// tree: parity unpinned for that formula).
#include "common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {

// mean / rstd of the pre-emphasised waveform, one CTA per utterance (deterministic two-level fp64 reduction)
__global__ void __launch_bounds__(256) preemph_stats_kernel(const float* __restrict__ x, long T, double eps,
                                                            float* __restrict__ mean_rstd) {
    __shared__ double scratch[32];
    const float* xb = x + (long)blockIdx.x * T;
    double s = 0.0, q = 0.0;
    for (long t = threadIdx.x; t < T; t += 256) {
        const float prev = t > 0 ? xb[t - 1] : xb[1];          // reflect padding of one sample on the left
        const float y = xb[t] - 0.97f * prev;
        s += (double)y; q += (double)y * (double)y;
    }
    s = block_sum(s, scratch);
    q = block_sum(q, scratch);
    if (threadIdx.x == 0) {
        const double mean = s / (double)T;
        double var = q / (double)T - mean * mean;
        if (var < 0.0) var = 0.0;
        mean_rstd[2 * blockIdx.x] = (float)mean;
        mean_rstd[2 * blockIdx.x + 1] = (float)(1.0 / sqrt(var + eps));
    }
}

// filtT[j][f], f < 2*nb: band-pass pair (cos = even, sin = odd) of band f % nb; kernel = 2*half + 1 taps
__global__ void sinc_filters_kernel(const float* __restrict__ low_hz_, const float* __restrict__ band_hz_,
                                    const float* __restrict__ window_, const float* __restrict__ n_, int nb, int half,
                                    float sample_rate, float min_low_hz, float min_band_hz, float* __restrict__ filtT) {
    const int kernel = 2 * half + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kernel * 2 * nb) return;
    const int j = idx / (2 * nb), f = idx % (2 * nb), band_i = f % nb;
    const bool is_sin = f >= nb;
    const float low = min_low_hz + fabsf(low_hz_[band_i]);
    float high = low + min_band_hz + fabsf(band_hz_[band_i]);
    high = fminf(fmaxf(high, min_low_hz), sample_rate * 0.5f);
    const float band = high - low;
    float v;
    if (j == half) {
        v = is_sin ? 0.f : 2.f * band;
    } else {
        const int jj = j < half ? j : kernel - 1 - j;            // mirrored tap of the left half
        const float n = n_[jj], w = window_[jj];
        const float left = is_sin ? (cosf(low * n) - cosf(high * n)) / (n * 0.5f) * w
                                  : (sinf(high * n) - sinf(low * n)) / (n * 0.5f) * w;
        v = (j < half) ? left : (is_sin ? -left : left);
    }
    filtT[idx] = v / (2.f * band);
}

// out[b, t', f] = log(|sum_j filtT[j][f] * yn[b, stride*t' + j]| + 1e-6), yn = instance-normalised pre-emphasised x.
// One CTA = TF frames x NF (= blockDim) filters; the signal window sits in shared memory, every filter tap is loaded
// once per CTA (coalesced over f) and reused for the TF frames held in registers.
template <int TF>
__global__ void __launch_bounds__(256) sinc_frontend_kernel(const float* __restrict__ x, long T, long Tp, int stride,
                                                            int kernel, const float* __restrict__ mean_rstd,
                                                            const float* __restrict__ in_w, const float* __restrict__ in_b,
                                                            const float* __restrict__ filtT, int nf,
                                                            float* __restrict__ out) {
    extern __shared__ float sig[];                       // (TF-1)*stride + kernel samples
    const int b = blockIdx.y;
    const long t0 = (long)blockIdx.x * TF;
    const float* xb = x + (long)b * T;
    const float mean = mean_rstd[2 * b], rstd = mean_rstd[2 * b + 1], gw = in_w[0], gb = in_b[0];
    const int nsig = (TF - 1) * stride + kernel;
    for (int i = threadIdx.x; i < nsig; i += blockDim.x) {
        const long t = t0 * stride + i;
        float v = 0.f;
        if (t < T) {
            const float prev = t > 0 ? xb[t - 1] : xb[1];
            v = ((xb[t] - 0.97f * prev) - mean) * rstd * gw + gb;
        }
        sig[i] = v;
    }
    __syncthreads();
    for (int f = threadIdx.x; f < nf; f += blockDim.x) {
        float acc[TF];
#pragma unroll
        for (int i = 0; i < TF; ++i) acc[i] = 0.f;
        for (int j = 0; j < kernel; ++j) {
            const float w = __ldg(filtT + (long)j * nf + f);
#pragma unroll
            for (int i = 0; i < TF; ++i) acc[i] = fmaf(w, sig[i * stride + j], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < TF; ++i)
            if (t0 + i < Tp) out[((long)b * Tp + t0 + i) * nf + f] = logf(fabsf(acc[i]) + 1e-6f);
    }
}

// x[b, t, c] -= mean_t x[b, t, c]; one CTA per (utterance, 32-channel group), fixed reduction order
__global__ void __launch_bounds__(256) time_mean_sub_kernel(float* __restrict__ x, long Tp, int C) {
    __shared__ double sh[8][32];
    __shared__ float smean[32];
    const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    float* xb = x + (long)b * Tp * C;
    double s = 0.0;
    if (c < C) for (long t = r; t < Tp; t += 8) s += (double)xb[t * C + c];
    sh[r][threadIdx.x & 31] = s;
    __syncthreads();
    if (r == 0) {
        for (int i = 1; i < 8; ++i) s += sh[i][threadIdx.x & 31];
        smean[threadIdx.x & 31] = (float)(s / (double)Tp);
    }
    __syncthreads();
    const float m = smean[threadIdx.x & 31];
    if (c < C) for (long t = r; t < Tp; t += 8) xb[t * C + c] -= m;
}

// Polyphase FIR resampling of the reference utterance (torchaudio.transforms.Resample as the RawNet inferencer / trainer
// apply it, src/inferencers/inferencer_rawnet.py:21-23,36): out[b, q*nw + i] = sum_k kern[i, k] * x[b, q*orig + k - width]
// (zero outside the signal), taps = 2*width + orig.
__global__ void __launch_bounds__(256) resample_fir_kernel(const float* __restrict__ x, const float* __restrict__ kern,
                                                           float* __restrict__ out, long T, long To, int orig, int nw,
                                                           int taps, int width, long total) {
    extern __shared__ float ks[];
    for (int i = threadIdx.x; i < nw * taps; i += blockDim.x) ks[i] = kern[i];
    __syncthreads();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long b = e / To, n = e - b * To;
        const long q = n / nw;
        const int i = (int)(n - q * nw);
        const float* xb = x + b * T;
        const long base = q * orig - width;
        const float* kk = ks + i * taps;
        float acc = 0.f;
        for (int k = 0; k < taps; ++k) {
            const long j = base + k;
            if (j >= 0 && j < T) acc = fmaf(kk[k], __ldg(xb + j), acc);
        }
        out[e] = acc;
    }
}

}  // namespace dprnn

using namespace dprnn;

extern "C" int dprnn_rawnet_frontend(const float* wave, int B, long T, const float* in_w, const float* in_b,
                                     const float* low_hz, const float* band_hz, const float* window, const float* n_half,
                                     int n_filters, int kernel, int stride, float sample_rate, float* filt_scratch,
                                     float* stats_scratch, float* out, void* stream) {
    DPRNN_CHECK_ARG(wave && in_w && in_b && low_hz && band_hz && window && n_half && filt_scratch && stats_scratch && out);
    DPRNN_CHECK_ARG(B > 0 && B <= 65535 && T >= kernel && T >= 2 && n_filters > 0 && n_filters % 2 == 0 && kernel % 2 == 1 && stride > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const long Tp = (T - kernel) / stride + 1;
    preemph_stats_kernel<<<B, 256, 0, st>>>(wave, T, 1e-4, stats_scratch);
    DPRNN_CHECK_LAUNCH();
    sinc_filters_kernel<<<cdiv((long)kernel * n_filters, 256), 256, 0, st>>>(low_hz, band_hz, window, n_half, n_filters / 2,
                                                                             kernel / 2, sample_rate, 50.f, 50.f, filt_scratch);
    DPRNN_CHECK_LAUNCH();
    constexpr int TF = 16;
    const size_t smem = ((TF - 1) * (size_t)stride + kernel) * sizeof(float);
    dim3 grid(cdiv(Tp, TF), B);
    sinc_frontend_kernel<TF><<<grid, 256, smem, st>>>(wave, T, Tp, stride, kernel, stats_scratch, in_w, in_b, filt_scratch,
                                                      n_filters, out);
    DPRNN_CHECK_LAUNCH();
    dim3 g2(cdiv(n_filters, 32), B);
    time_mean_sub_kernel<<<g2, 256, 0, st>>>(out, Tp, n_filters);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Res2Net blocks + attentive statistics pooling: the stages around the contractions (which run in gemm_tc.cu)
// ---------------------------------------------------------------------------------------------------------
namespace dprnn {

static inline unsigned rn_grid(long total, int threads) {
    long g = (total + threads - 1) / threads;
    const long cap = 148L * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

__global__ void res2_gather_kernel(const float* __restrict__ a, long lda, const float* __restrict__ b, long ldb,
                                   float* __restrict__ col, long rows, long T, int c4n, int dil) {
    const long total = rows * 3 * c4n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        const long rt = idx / c4n;
        const int tap = (int)(rt % 3);
        const long r = rt / 3;
        const long t = r % T, ts = t + (long)(tap - 1) * dil;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ts >= 0 && ts < T) {
            const long rs = r + (long)(tap - 1) * dil;
            v = *reinterpret_cast<const float4*>(a + rs * lda + 4 * c4);
            if (b) {
                const float4 w = *reinterpret_cast<const float4*>(b + rs * ldb + 4 * c4);
                v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            }
        }
        reinterpret_cast<float4*>(col)[idx] = v;          // col[r][tap*C + c]
    }
}

__global__ void maxpool_time_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                                    long ldo, int B, long T, long To, int c4n, int k) {
    const long total = (long)B * To * c4n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        const long r = idx / c4n;
        const long to = r % To, b = r / To;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int i = 0; i < k; ++i) {
            const long src = ((b * T + to * k + i) * c4n + c4);
            float4 v = reinterpret_cast<const float4*>(x)[src];
            if (y) {
                const float4 w = reinterpret_cast<const float4*>(y)[src];
                v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            }
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
        *reinterpret_cast<float4*>(out + r * ldo + 4 * c4) = m;
    }
}

// one CTA per (utterance, 32-channel group): mean (and clamped unbiased std) over time, fixed reduction order
__global__ void __launch_bounds__(256) col_mean_std_kernel(const float* __restrict__ x, float* __restrict__ mean,
                                                           float* __restrict__ stdv, long T, int C) {
    __shared__ double sh[2][8][32];
    const int b = blockIdx.y, cl = threadIdx.x & 31, c = blockIdx.x * 32 + cl, r = threadIdx.x >> 5;
    const float* xb = x + (long)b * T * C;
    double s = 0.0, q = 0.0;
    if (c < C) for (long t = r; t < T; t += 8) { const double v = xb[t * C + c]; s += v; q += v * v; }
    sh[0][r][cl] = s; sh[1][r][cl] = q;
    __syncthreads();
    if (r == 0 && c < C) {
        for (int i = 1; i < 8; ++i) { s += sh[0][i][cl]; q += sh[1][i][cl]; }
        const double m = s / (double)T;
        mean[(long)b * C + c] = (float)m;
        if (stdv) {
            double var = T > 1 ? (q - (double)T * m * m) / (double)(T - 1) : 0.0;      // torch.var: unbiased
            var = fmin(fmax(var, 1e-4), 1e4);
            stdv[(long)b * C + c] = (float)sqrt(var);
        }
    }
}

__global__ void afms_apply_kernel(const float* __restrict__ x, const float* __restrict__ alpha,
                                  const float* __restrict__ gate, float* __restrict__ out, long ldo, long T, long total4,
                                  int c4n) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        const long r = idx / c4n, b = r / T;
        const float4 v = reinterpret_cast<const float4*>(x)[idx];
        const float4 al = __ldg(reinterpret_cast<const float4*>(alpha) + c4);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gate) + b * c4n + c4);
        *reinterpret_cast<float4*>(out + r * ldo + 4 * c4) =
            make_float4((v.x + al.x) * g.x, (v.y + al.y) * g.y, (v.z + al.z) * g.z, (v.w + al.w) * g.w);
    }
}

__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long total4) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total4; idx += (long)gridDim.x * blockDim.x) {
        const float4 u = reinterpret_cast<const float4*>(a)[idx], v = reinterpret_cast<const float4*>(b)[idx];
        reinterpret_cast<float4*>(out)[idx] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    }
}

__global__ void affine_vec_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                  const float* __restrict__ shift, float* __restrict__ out, int total, int C, int act) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = idx % C;
    float v = x[idx];
    if (scale) v = fmaf(v, scale[c], shift[c]);
    if (act == 1) v = sigmoid_acc(v);
    out[idx] = v;
}

// one thread per (utterance, channel): softmax over time of the logits, weighted mean / std of x
__global__ void att_stats_pool_kernel(const float* __restrict__ x, const float* __restrict__ logits,
                                      float* __restrict__ out, int B, long T, int C) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * C) return;
    const int b = idx / C, c = idx % C;
    const float* xb = x + (long)b * T * C + c;
    const float* lb = logits + (long)b * T * C + c;
    float m = -INFINITY;
    for (long t = 0; t < T; ++t) m = fmaxf(m, lb[t * C]);
    double se = 0.0, sx = 0.0, sxx = 0.0;
    for (long t = 0; t < T; ++t) {
        const double e = (double)expf(lb[t * C] - m), v = xb[t * C];
        se += e; sx += e * v; sxx += e * v * v;
    }
    const double mu = sx / se;
    double var = sxx / se - mu * mu;
    var = fmin(fmax(var, 1e-4), 1e4);
    out[(long)b * 2 * C + c] = (float)mu;
    out[(long)b * 2 * C + C + c] = (float)sqrt(var);
}

}  // namespace dprnn

extern "C" int dprnn_res2_gather(const float* a, long lda, const float* b, long ldb, float* col, long rows, long T, int C,
                                 int dil, void* stream) {
    DPRNN_CHECK_ARG(a && col && rows > 0 && T > 0 && rows % T == 0 && C % 4 == 0 && dil > 0 && lda % 4 == 0 && (!b || ldb % 4 == 0));
    res2_gather_kernel<<<rn_grid(rows * 3 * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, col, rows, T,
                                                                                          C / 4, dil);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_maxpool_time(const float* x, const float* y, float* out, long ldo, int B, long T, int C, int k,
                                  void* stream) {
    DPRNN_CHECK_ARG(x && out && B > 0 && T >= k && k > 0 && C % 4 == 0 && ldo % 4 == 0);
    const long To = T / k;
    maxpool_time_kernel<<<rn_grid((long)B * To * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, y, out, ldo, B, T, To,
                                                                                               C / 4, k);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_col_mean_std(const float* x, float* mean, float* stdv, int B, long T, int C, void* stream) {
    DPRNN_CHECK_ARG(x && mean && B > 0 && B <= 65535 && T > 0 && C > 0);
    dim3 grid(cdiv(C, 32), B);
    col_mean_std_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, mean, stdv, T, C);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_afms_apply(const float* x, const float* alpha, const float* gate, float* out, long ldo, int B, long T,
                                int C, void* stream) {
    DPRNN_CHECK_ARG(x && alpha && gate && out && B > 0 && T > 0 && C % 4 == 0 && ldo % 4 == 0);
    afms_apply_kernel<<<rn_grid((long)B * T * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, alpha, gate, out, ldo, T,
                                                                                            (long)B * T * (C / 4), C / 4);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_add2(const float* a, const float* b, float* out, long n, void* stream) {
    DPRNN_CHECK_ARG(a && b && out && n > 0 && n % 4 == 0);
    add2_kernel<<<rn_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n / 4);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_affine_vec(const float* x, const float* scale, const float* shift, float* out, int B, int C, int act,
                                void* stream) {
    DPRNN_CHECK_ARG(x && out && B > 0 && C > 0 && ((scale == nullptr) == (shift == nullptr)));
    affine_vec_kernel<<<cdiv((long)B * C, 256), 256, 0, (cudaStream_t)stream>>>(x, scale, shift, out, B * C, C, act);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_att_stats_pool(const float* x, const float* logits, float* out, int B, long T, int C, void* stream) {
    DPRNN_CHECK_ARG(x && logits && out && B > 0 && T > 0 && C > 0);
    att_stats_pool_kernel<<<cdiv((long)B * C, 128), 128, 0, (cudaStream_t)stream>>>(x, logits, out, B, T, C);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_resample_fir(const float* x, const float* kernel, float* out, int B, long T, long To, int orig, int nw,
                                  int taps, int width, void* stream) {
    DPRNN_CHECK_ARG(x && kernel && out && B > 0 && T > 0 && To > 0 && orig > 0 && nw > 0 && taps > 0 && width >= 0);
    DPRNN_CHECK_ARG((size_t)nw * taps * sizeof(float) <= 48 * 1024);
    const long total = (long)B * To;
    const long want = (total + 255) / 256;
    const unsigned grid = (unsigned)(want < 148L * 16 ? want : 148L * 16);
    resample_fir_kernel<<<grid, 256, (size_t)nw * taps * sizeof(float), (cudaStream_t)stream>>>(
        x, kernel, out, T, To, orig, nw, taps, width, total);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
