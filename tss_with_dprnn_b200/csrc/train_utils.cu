// The callers' side of the path: the evaluation metric and the optimiser step that the
// reference's drivers run around the forward, as device kernels so that a batched evaluation / data-parallel training
// step never leaves the GPU:
//   * SI-SDR per utterance (asteroid's pairwise_neg_sisdr / get_metrics('si_sdr') recipe used by
//     src/trainers/trainer_spe.py:39 and src/inferencers/inferencer_spe.py:37-43): zero-mean both signals,
//     s = <e,t> t / (|t|^2 + eps), 10 log10(|s|^2 / (|e - s|^2 + eps) + eps), eps = 1e-8;
//   * clip_grad_norm_(params, max_norm) + Adam(lr, betas, eps, weight_decay) over ONE flat fp32 buffer
//     (src/trainers/trainer.py:42-43,115-116; scripts/train/config_tss.yaml:36-39,59): the global norm is reduced
//     deterministically, the update is a single pass.
//   * the TrainerSpe loss (src/trainers/trainer_spe.py:39-43): mean negative SI-SDR + ce_gamma * CrossEntropy(logits, spk),
//     forward value and the gradients w.r.t. est / logits in one launch (the backward of the path starts from them).
#include "common.cuh"
#include "../../include/dprnn_b200.h"

namespace dprnn {

// one CTA per utterance; signals est / target at offsets off[b] (or b * len) with len[b] samples
__global__ void __launch_bounds__(256) si_sdr_kernel(const float* __restrict__ est, const float* __restrict__ tgt,
                                                     const long* __restrict__ off, const long* __restrict__ len,
                                                     long uniform_len, float* __restrict__ out_db) {
    __shared__ double scratch[32];
    const int b = blockIdx.x;
    const long o = off ? off[b] : (long)b * uniform_len, n = len ? len[b] : uniform_len;
    const float* e = est + o;
    const float* t = tgt + o;
    double se = 0, st = 0, see = 0, stt = 0, set = 0;
    for (long i = threadIdx.x; i < n; i += 256) {
        const double a = e[i], c = t[i];
        se += a; st += c; see += a * a; stt += c * c; set += a * c;
    }
    se = block_sum(se, scratch); st = block_sum(st, scratch); see = block_sum(see, scratch);
    stt = block_sum(stt, scratch); set = block_sum(set, scratch);
    if (threadIdx.x == 0) {
        const double N = (double)n, eps = 1e-8;
        const double ee = see - se * se / N, tt = stt - st * st / N, et = set - se * st / N;   // zero-mean moments
        const double alpha = et / (tt + eps);
        const double s2 = alpha * alpha * tt;                  // |s|^2
        const double n2 = ee - 2.0 * alpha * et + s2;          // |e - s|^2
        out_db[b] = (float)(10.0 * log10(s2 / (fmax(n2, 0.0) + eps) + eps));
    }
}

// loss_b = -SI-SDR(est_b, tgt_b) + gamma * CE(logits_b, spk_b); one CTA per utterance.  Writes the two terms to
// terms[b*2 + {0,1}] and the gradients of  (1/B) sum_b loss_b  to d_est / d_logits.
// With e', t' the zero-mean signals, a = |s|^2, n = |e' - s|^2, r = a / (n + eps):
//   dr/de' = (2 alpha tt/(tt+eps)) t' / (n+eps) - a/(n+eps)^2 * (2 (e' - alpha t') - 2 (et - alpha tt)/(tt+eps) t')
// and dL/dr = -10 / (ln 10 (r + eps)).  e' and t' are zero-mean, so the mean-removal adjoint is the identity on them.
__global__ void __launch_bounds__(256) train_loss_kernel(const float* __restrict__ est, const float* __restrict__ tgt,
                                                         long T, const float* __restrict__ logits, int C,
                                                         const long* __restrict__ spk, float gamma, int B,
                                                         float* __restrict__ terms, float* __restrict__ d_est,
                                                         float* __restrict__ d_logits) {
    __shared__ double scratch[32];
    __shared__ double coef[4];
    const int b = blockIdx.x;
    const float* e = est + (long)b * T;
    const float* t = tgt + (long)b * T;
    double se = 0, st = 0, see = 0, stt = 0, set = 0;
    for (long i = threadIdx.x; i < T; i += 256) {
        const double a = e[i], c = t[i];
        se += a; st += c; see += a * a; stt += c * c; set += a * c;
    }
    se = block_sum(se, scratch); st = block_sum(st, scratch); see = block_sum(see, scratch);
    stt = block_sum(stt, scratch); set = block_sum(set, scratch);
    if (threadIdx.x == 0) {
        const double N = (double)T, eps = 1e-8;
        const double ee = see - se * se / N, tt = stt - st * st / N, et = set - se * st / N;
        const double alpha = et / (tt + eps);
        const double a2 = alpha * alpha * tt;
        const double n2 = fmax(ee - 2.0 * alpha * et + a2, 0.0);
        const double r = a2 / (n2 + eps);
        terms[b * 2] = (float)(-10.0 * log10(r + eps));
        const double dLdr = -10.0 / (log(10.0) * (r + eps)) / (double)B;
        const double proj = (et - alpha * tt) / (tt + eps);                 // <e' - alpha t', t'> / (tt + eps)
        // d r / d e' = 2 alpha tt/((tt+eps)(n+eps)) t' - a/(n+eps)^2 * 2 (e' - alpha t' - proj t')
        const double ce_ = -2.0 * a2 / ((n2 + eps) * (n2 + eps));           // coefficient of e'
        const double ct_full = 2.0 * alpha * tt / (tt + eps) / (n2 + eps) - ce_ * (alpha + proj);
        coef[0] = dLdr * ce_; coef[1] = dLdr * ct_full; coef[2] = se / N; coef[3] = st / N;
    }
    __syncthreads();
    const double c_e = coef[0], c_t = coef[1], me = coef[2], mt = coef[3];
    for (long i = threadIdx.x; i < T; i += 256)
        d_est[(long)b * T + i] = (float)(c_e * ((double)e[i] - me) + c_t * ((double)t[i] - mt));
    if (!logits) {                       // BSS trainer: SI-SDR only
        if (threadIdx.x == 0) terms[b * 2 + 1] = 0.f;
        return;
    }
    // cross entropy over C classes
    const float* lg = logits + (long)b * C;
    double mx = -1e300;
    for (int j = threadIdx.x; j < C; j += 256) mx = fmax(mx, (double)lg[j]);
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = scratch[0];
    for (int w = 1; w < 8; ++w) mx = fmax(mx, scratch[w]);
    double sum = 0.0;
    for (int j = threadIdx.x; j < C; j += 256) sum += exp((double)lg[j] - mx);
    sum = block_sum(sum, scratch);
    const long y = spk[b];
    if (threadIdx.x == 0) terms[b * 2 + 1] = (float)(gamma * (mx + log(sum) - (double)lg[y]));
    for (int j = threadIdx.x; j < C; j += 256)
        d_logits[(long)b * C + j] = (float)((double)gamma / B * (exp((double)lg[j] - mx) / sum - (j == y ? 1.0 : 0.0)));
}

// Permutation-invariant assignment for two sources (asteroid PITLossWrapper(pairwise_neg_sisdr, pit_from='pw_mtx') as
// src/trainers/trainer.py:39 builds it): per utterance the pairwise neg-SI-SDR matrix of est [B,2,T] x target [B,2,T],
// the permutation with the smaller summed loss (ties: identity, the first in asteroid's permutation list), and the
// targets re-ordered accordingly (target_perm [B,2,T]) so that the loss / gradient kernel runs on matched rows.
__global__ void __launch_bounds__(256) pit2_kernel(const float* __restrict__ est, const float* __restrict__ tgt, long T,
                                                   float* __restrict__ tgt_perm, int* __restrict__ perm,
                                                   float* __restrict__ pw) {
    __shared__ double scratch[32];
    __shared__ int choice;
    const int b = blockIdx.x;
    const float* e0 = est + (long)b * 2 * T; const float* e1 = e0 + T;
    const float* t0 = tgt + (long)b * 2 * T; const float* t1 = t0 + T;
    double s[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = 0.0;
    for (long i = threadIdx.x; i < T; i += 256) {
        const double a0 = e0[i], a1 = e1[i], c0 = t0[i], c1 = t1[i];
        s[0] += a0; s[1] += a1; s[2] += c0; s[3] += c1;
        s[4] += a0 * a0; s[5] += a1 * a1; s[6] += c0 * c0; s[7] += c1 * c1;
        s[8] += a0 * c0; s[9] += a0 * c1; s[10] += a1 * c0; s[11] += a1 * c1;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = block_sum(s[i], scratch);
    if (threadIdx.x == 0) {
        const double N = (double)T, eps = 1e-8;
        double L[2][2];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                const double ee = s[4 + i] - s[i] * s[i] / N, tt = s[6 + j] - s[2 + j] * s[2 + j] / N;
                const double et = s[8 + 2 * i + j] - s[i] * s[2 + j] / N;
                const double alpha = et / (tt + eps), a2 = alpha * alpha * tt;
                const double n2 = fmax(ee - 2.0 * alpha * et + a2, 0.0);
                L[i][j] = -10.0 * log10(a2 / (n2 + eps) + eps);
                if (pw) pw[b * 4 + i * 2 + j] = (float)L[i][j];
            }
        choice = (L[0][1] + L[1][0] < L[0][0] + L[1][1]) ? 1 : 0;
        perm[b] = choice;
    }
    __syncthreads();
    const float* src0 = choice ? t1 : t0;
    const float* src1 = choice ? t0 : t1;
    float* d0 = tgt_perm + (long)b * 2 * T;
    for (long i = threadIdx.x; i < T; i += 256) { d0[i] = src0[i]; d0[T + i] = src1[i]; }
}

__global__ void train_loss_final_kernel(const float* __restrict__ terms, int B, float* __restrict__ loss) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < B; i += 32) { s0 += terms[i * 2]; s1 += terms[i * 2 + 1]; }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if (threadIdx.x == 0) { loss[0] = (float)((s0 + s1) / B); loss[1] = (float)(s0 / B); loss[2] = (float)(s1 / B); }
}

// 16-bit PCM -> float32 as libsndfile / soundfile.read(dtype='float32') normalises it (x / 32768, exact): the packed
// shards of shards.py travel to the device as int16 (half the H2D bytes) and are widened here
__global__ void pcm16_to_f32_kernel(const short* __restrict__ in, float* __restrict__ out, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = (float)in[i] * (1.0f / 32768.0f);
}

__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const float* __restrict__ g, long n, double* __restrict__ partial) {
    __shared__ double scratch[32];
    const long per = (n + gridDim.x - 1) / gridDim.x;
    const long beg = (long)blockIdx.x * per, end = min(beg + per, n);
    double s = 0.0;
    for (long i = beg + threadIdx.x; i < end; i += 256) { const double v = g[i]; s += v * v; }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void sqnorm_final_kernel(const double* __restrict__ partial, int nparts, float* __restrict__ total_norm) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 32) s += partial[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) total_norm[0] = (float)sqrt(s);
}

// torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (L2 weight decay added to the gradient, bias-corrected moments)
__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long n, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, float max_norm, float bc1, float bc2_sqrt,
                                 const float* __restrict__ total_norm) {
    float clip = 1.0f;
    if (max_norm > 0.f) {
        clip = max_norm / (total_norm[0] + 1e-6f);
        clip = clip < 1.0f ? clip : 1.0f;
    }
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float gi = g[i] * clip;
        const float pi = p[i];
        gi = fmaf(weight_decay, pi, gi);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

}  // namespace dprnn

using namespace dprnn;

extern "C" int dprnn_si_sdr(const float* est, const float* target, const long* off, const long* len, long uniform_len,
                            int B, float* out_db, void* stream) {
    DPRNN_CHECK_ARG(est && target && out_db && B > 0 && ((off && len) || uniform_len > 0));
    si_sdr_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(est, target, off, len, uniform_len, out_db);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_train_loss(const float* est, const float* target, long T, const float* logits, int C, const long* spk,
                                float ce_gamma, int B, float* terms, float* loss3, float* d_est, float* d_logits,
                                void* stream) {
    DPRNN_CHECK_ARG(est && target && terms && loss3 && d_est && B > 0 && T > 0);
    DPRNN_CHECK_ARG(!logits || (spk && d_logits && C > 0));
    train_loss_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(est, target, T, logits, C, spk, ce_gamma, B, terms, d_est,
                                                           d_logits);
    DPRNN_CHECK_LAUNCH();
    train_loss_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(terms, B, loss3);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_pit2_assign(const float* est, const float* target, int B, long T, float* target_perm, int* perm,
                                 float* pairwise, void* stream) {
    DPRNN_CHECK_ARG(est && target && target_perm && perm && B > 0 && T > 0);
    pit2_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(est, target, T, target_perm, perm, pairwise);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" int dprnn_pcm16_to_f32(const void* pcm, float* out, long n, void* stream) {
    DPRNN_CHECK_ARG(pcm && out && n > 0);
    const long want = (n + 255) / 256;
    pcm16_to_f32_kernel<<<(unsigned)(want < 148L * 16 ? want : 148L * 16), 256, 0, (cudaStream_t)stream>>>(
        (const short*)pcm, out, n);
    DPRNN_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t dprnn_clip_adam_workspace_bytes(void) { return 1024 * sizeof(double) + 256; }

extern "C" int dprnn_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, float lr,
                                    float beta1, float beta2, float eps, float weight_decay, float max_norm, int step,
                                    void* workspace, float* total_norm_out, void* stream) {
    DPRNN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && workspace && total_norm_out && n > 0 && step >= 1);
    cudaStream_t st = (cudaStream_t)stream;
    const int nparts = 1024;
    sqnorm_partial_kernel<<<nparts, 256, 0, st>>>(grads, n, (double*)workspace);
    DPRNN_CHECK_LAUNCH();
    sqnorm_final_kernel<<<1, 32, 0, st>>>((const double*)workspace, nparts, total_norm_out);
    DPRNN_CHECK_LAUNCH();
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    const unsigned grid = (unsigned)((n + 255) / 256 < 148L * 16 ? (n + 255) / 256 : 148L * 16);
    clip_adam_kernel<<<grid, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                           max_norm, bc1, bc2_sqrt, total_norm_out);
    DPRNN_CHECK_LAUNCH();
    return 0;
}
