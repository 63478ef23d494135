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
