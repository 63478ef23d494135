// The callers' side of the path: the evaluation metric and the optimiser step that the
// reference's drivers run around the forward, as device kernels so that a batched evaluation / data-parallel training
// step never leaves the GPU:
//   * SI-SDR per utterance (asteroid's pairwise_neg_sisdr / get_metrics('si_sdr') recipe used by
//     src/trainers/trainer_spe.py:39 and src/inferencers/inferencer_spe.py:37-43): zero-mean both signals,
//     s = <e,t> t / (|t|^2 + eps), 10 log10(|s|^2 / (|e - s|^2 + eps) + eps), eps = 1e-8;
//   * clip_grad_norm_(params, max_norm) + Adam(lr, betas, eps, weight_decay) over ONE flat fp32 buffer
//     (src/trainers/trainer.py:42-43,115-116; scripts/train/config_tss.yaml:36-39,59): the global norm is reduced
//     deterministically, the update is a single pass.
// The backward kernels that would fill the gradient buffer are not built yet (cfg 5).
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
