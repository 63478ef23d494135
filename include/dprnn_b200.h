/* libdprnn_b200 - C ABI of the B200-native DPRNN separation forward path.
 *
 * Drop-in boundary for the reference's `src/models` forward (Aleksashka-i/tss-with-dprnn).  The
 * reference has no FFI of its own - its boundary is the nn.Module API - so
 * each entry point below names the reference statement(s) it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - activations are fp32, channels-last: [B, L, C] frames / [B, S, K, F] chunks (the reference
 *     uses [B, C, L] / [B, F, K, S]); weights that feed a contraction are passed TRANSPOSED
 *     ([K_in, N_out] row-major) - the host side packs them once per forward from the state_dict;
 *   - the library never allocates: scratch is passed in, sizes come from *_workspace_bytes();
 *   - `stream` is a cudaStream_t; all work is enqueued on it and nothing synchronises;
 *   - return value 0 = ok, 1 = CUDA error, 2 = bad argument; dprnn_last_error() has the text.
 */
#ifndef DPRNN_B200_H
#define DPRNN_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* dprnn_last_error(void);
/* "arch=sm_100a;..." - which kernels / precision modes this build carries. */
const char* dprnn_build_info(void);

/* epilogues of dprnn_gemm_* */
enum { DPRNN_EPI_NONE = 0, DPRNN_EPI_RELU = 1, DPRNN_EPI_SIGMOID = 2, DPRNN_EPI_GATED = 3, DPRNN_EPI_AFFINE_PRELU = 4,
       DPRNN_EPI_RELU_AFFINE = 5 };

/* Encoder.forward, src/models/encoder_decoder.py:25-33 (Conv1d(1->N, ksz, stride, bias=False) + ReLU).
 * wave [B,T], w [N,ksz] -> enc [B, L=(T-ksz)/stride+1, N]. */
int dprnn_encoder_fwd(const float* wave, const float* w, float* enc, int B, int T, int N, int ksz, int stride,
                      void* stream);

/* Statistics of nn.GroupNorm(1,C) / norms.GlobLN (src/models/dprnn.py:72-77,130-134, norms.py:6-31):
 * per-utterance mean and 1/sqrt(biased var + eps) over a contiguous slab of elems_per_utt floats.
 * mean_rstd [B,2]. Deterministic two-level fp64 reduction. */
size_t dprnn_utt_stats_workspace_bytes(int B);
int dprnn_utt_stats(const float* x, int B, long elems_per_utt, float eps, void* workspace, float* mean_rstd,
                    void* stream);

/* Folds the norm's affine (and an optional per-utterance channel multiplier - the mul / FiLM / att
 * fusions of src/models/dprnn_spe.py:199-229) into s1[b,c] = gamma*rstd*mul, s0[b,c] = (beta-mean*gamma*rstd)*mul
 * for use as the GEMM prologue. mulc may be NULL. */
int dprnn_norm_affine(const float* mean_rstd, const float* gamma, const float* beta, const float* mulc, float* s1,
                      float* s0, int B, int C, void* stream);

/* intra_norm / inter_norm + residual, src/models/dprnn.py:90-92,98-99:
 * x[b,r,c] += (y[b,r,c]-mean_b)*rstd_b*gamma_c + beta_c.  x_bf16 (may be NULL): bf16 copy of the updated x,
 * the TMA-fed input of the next tensor-core LSTM layer. */
int dprnn_norm_residual(const float* y, float* x, const float* mean_rstd, const float* gamma, const float* beta,
                        int B, long rows_per_utt, int C, void* x_bf16, void* stream);
/* Out of place: x_out = x + norm(y), x untouched (the training backward walks the reversible residual stream back with
 * negated gamma / beta while a side stream still reads the layer's x_out: no wait between the two). */
int dprnn_norm_residual_to(const float* y, const float* x, const float* mean_rstd, const float* gamma, const float* beta,
                           int B, long rows_per_utt, int C, float* x_out, void* x_bf16, void* stream);

/* 16-bit storage / tensor-core operand formats (argument `h16` below).  bf16: fp32's range, 8 significand bits - the
 * 'bf16' mode.  fp16: 11 significand bits - the 'fp16' mode, which keeps the estimated sources within north_star's
 * 1e-3 of the reference's fp32 path at the same kernel speed.  Entry points named *_bf16* without an
 * `h16` argument are the bf16 instances of the *_h16* ones and are kept for existing callers. */
#define DPRNN_H16_BF16 0
#define DPRNN_H16_FP16 1

/* Same with y stored in 16 bits (the output of dprnn_linear_h16out_stats); C a multiple of 8; x_h16 may be NULL. */
int dprnn_norm_residual_yh16(const void* y_h16, float* x, const float* mean_rstd, const float* gamma, const float* beta,
                             int B, long rows_per_utt, int C, void* x_h16, int h16, void* stream);
int dprnn_norm_residual_ybf16(const void* y_bf16, float* x, const float* mean_rstd, const float* gamma, const float* beta,
                              int B, long rows_per_utt, int C, void* x_bf16, void* stream);
/* Opt-in variant with the residual stream kept in bf16 only: x_bf16 <- bf16(float(x_bf16) + norm(y)); when x_f32_out is
 * given (last half-block) the fp32 result is written there instead, for the fold.  Halves the HBM traffic of the stage at
 * the price of a bf16 rounding of the residual per half-block (Engine.residual_bf16, off by default). */
int dprnn_norm_residual_bf16res(const void* y_bf16, void* x_bf16, float* x_f32_out, const float* mean_rstd,
                                const float* gamma, const float* beta, int B, long rows_per_utt, int C, void* stream);
int dprnn_norm_residual_h16res(const void* y_h16, void* x_h16, float* x_f32_out, const float* mean_rstd,
                               const float* gamma, const float* beta, int B, long rows_per_utt, int C, int h16,
                               void* stream);

/* DPRNN._segmentation, src/models/dprnn.py:189-201 (F.unfold, kernel K, pad K, stride P):
 * y [B,L,F] -> x [B,S,K,F], x[b,s,k,:] = y[b, s*P+k-K, :] or 0. S = dprnn_num_chunks(L,K,P). Bit-exact. */
int dprnn_num_chunks(long L, int K, int P);
int dprnn_unfold(const float* y, float* x, int B, long L, int K, int P, int F, void* stream);
/* Same, also writing the bf16 copy x_bf16 [B,S,K,F] (operand of the first tensor-core LSTM layer); x may be NULL when the
 * residual stream is kept in 16 bits only (nothing reads the fp32 copy before the last half-block rewrites it). */
int dprnn_unfold_bf16(const float* y, float* x, void* x_bf16, int B, long L, int K, int P, int F, void* stream);
int dprnn_unfold_h16(const float* y, float* x, void* x_h16, int B, long L, int K, int P, int F, int h16, void* stream);

/* The last half-block's norm + residual (src/models/dprnn.py:98-99) + self.prelu + DPRNN._overlap_add (:174,203-217) in
 * one pass over the 16-bit residual stream: out[b,t,:] = sum over the (at most two) chunks s covering t of
 * prelu(float(x_h16[b,s,k,:]) + norm_b(float(y_h16[b,s,k,:]))), k = t + K - s*P.  y_h16 / x_h16 [B,S,K,F] in the format
 * h16, mean_rstd [B,2], out [B,L,F] fp32, F a multiple of 8.  Bit for bit dprnn_norm_residual_h16res (fp32 output) followed
 * by dprnn_fold_prelu, without the fp32 [B,S,K,F] tensor between them. */
int dprnn_norm_residual_fold_prelu_h16(const void* y_h16, const void* x_h16, const float* mean_rstd, const float* gamma,
                                       const float* beta, float* out, int B, long L, int K, int P, int F,
                                       const float* prelu_a, int h16, void* stream);

/* self.prelu + DPRNN._overlap_add, src/models/dprnn.py:174,203-217 (F.fold, plain sum):
 * x [B,S,K,F] -> out [B,L,F]. prelu_a (1 float, device) may be NULL for a pure fold. */
int dprnn_fold_prelu(const float* x, float* out, int B, long L, int K, int P, int F, const float* prelu_a,
                     void* stream);

/* masks * encoders then Decoder.forward, src/models/dprnn_spe.py:323-325, dprnn.py:277-281,
 * encoder_decoder.py:40-49 (ConvTranspose1d(N->1, ksz, stride, bias=False)).
 * mask: utterance b at mask + b*mask_utt_stride, [L,N]; enc [B,L,N]; wdec [N,ksz];
 * out: utterance b at out + b*out_utt_stride, T=(L-1)*stride+ksz samples. */
int dprnn_mask_decode(const float* mask, long mask_utt_stride, const float* enc, const float* wdec, float* out,
                      long out_utt_stride, int B, long L, int N, int ksz, int stride, void* stream);

/* d0 = (masks * input)[:,0], src/models/dprnn_spe_ira.py:79-80,107-108: out = mask * enc, elementwise. */
int dprnn_mask_apply(const float* mask, const float* enc, float* out, long elems, void* stream);

/* Attention fusion, src/models/dprnn_spe.py:177-183,217-225: depthwise average conv (stride ksz) of the
 * normalised encoding, channel dot with v = fusion_linear(e), softmax over time, nearest upsample to L.
 * Produces rowscale[b,l] = 1 + softmax[b, src(l)] (the per-channel factor v goes through dprnn_norm_affine).
 * scores [B,La] is scratch/out (La=(L-ksz)/ksz+1); the upsample index map is ATen's, bit-exact. */
int dprnn_att_rowscale(const float* enc, const float* s1, const float* s0, const float* wavg, const float* bavg,
                       const float* v, float* scores, float* rowscale, int B, long L, int N, int ksz, void* stream);

/* nn.BatchNorm1d of ResBlock, src/models/dprnn_spe.py:20-21,33,36: per-channel scale/shift.
 * training!=0: batch statistics of y [rows,C] + running-stat update in place (unbiased var, momentum);
 * training==0: running statistics (y, workspace unused). C must divide 256. */
size_t dprnn_bn_workspace_bytes(int C);
int dprnn_batchnorm_affine(const float* y, long rows, int C, const float* weight, const float* bias,
                           float* running_mean, float* running_var, int training, float eps, float momentum,
                           void* workspace, float* scale, float* shift, void* stream);

/* BN apply + PReLU, src/models/dprnn_spe.py:33-34. */
int dprnn_affine_prelu(const float* y, const float* scale, const float* shift, const float* prelu_a, float* out,
                       long rows, int C, void* stream);
/* BN apply + skip + PReLU + MaxPool1d(3), src/models/dprnn_spe.py:36-42. y, skip [B,Lin,C] -> out [B,Lin/3,C]. */
int dprnn_affine_add_prelu_pool3(const float* y, const float* scale, const float* shift, const float* skip,
                                 const float* prelu_a, float* out, int B, long Lin, int C, void* stream);

/* Time sum of the speaker encoder output divided by the per-utterance length div[b],
 * src/models/dprnn_spe.py:159-161. x [B,Lx,C] -> emb [B,C]. */
int dprnn_time_sum(const float* x, float* emb, int B, long Lx, int C, const float* div, void* stream);

/* Tiny Linear on embeddings (fusion_linear*, pred_linear, aux_linear; src/models/dprnn_spe.py:92-98,123,
 * dprnn_spe_ira.py:51): out[b,n] (+)= bias[n] + sum_k in[b*ldin+k]*W[n*ldw+k]. W in nn.Linear layout. */
int dprnn_small_linear(const float* in, long ldin, const float* W, long ldw, const float* bias, float* out,
                       long ldout, int B, int N, int K, int accumulate, void* stream);

/* Every pointwise contraction of the path in exact fp32 (1x1 Conv1d / Conv2d / nn.Linear:
 * src/models/dprnn.py:61,70,135,155,157-160; dprnn_spe.py:17-18,117,121):
 *   C[M,N] = epi( pro(A)[M,K] @ Wt[K,N] + bias_scale*bias )
 * pro: a = (a*p_scale[b,k] + p_shift[b,k]) * rowscale[row] + p_add[b,k] with b = row / rows_per_utt
 *      (any of p_scale/p_shift, rowscale, p_add may be NULL);
 * bias: [N], or [B,N] when bias_per_utt;
 * DPRNN_EPI_GATED: Wt/bias columns packed per 128-wide tile as 64 'out' + 64 'gate' units and
 *      C[M,N/2] = tanh(out) * sigmoid(gate)  (src/models/dprnn.py:181). */
int dprnn_gemm_f32(const float* A, long lda, const float* Wt, long ldw, float* C, long ldc, int M, int N, int K,
                   const float* bias, int bias_per_utt, float bias_scale, long rows_per_utt, const float* p_scale,
                   const float* p_shift, const float* p_add, const float* rowscale, int epilogue, void* stream);

/* The sequential half of nn.LSTM, src/models/dprnn.py:23-28,35-36 (gate order i,f,g,o, zero initial
 * state), exact fp32 on CUDA cores.  gx [rows, ndir*4H] = x W_ih^T + b_ih + b_hh (from dprnn_gemm_f32),
 * whhT [ndir][H][4H], hout [rows, ndir*H].  Sequence n, step t lives at row
 *   (n / seq_div)*seq_outer_stride + (n % seq_div)*seq_inner_stride + t*step_stride
 * (intra-chunk: seq_div=1, outer=K, step=1; inter-chunk: seq_div=K, outer=S*K, inner=1, step=K).
 * Direction 1 walks t = T-1..0. hidden must be 128. */
int dprnn_lstm_recurrence_f32(const float* gx, const float* whhT, float* hout, long nseq, int T, long seq_div,
                              long seq_outer_stride, long seq_inner_stride, long step_stride, int hidden, int ndir,
                              void* stream);

/* ---- bf16 tensor-core mode (tcgen05 + TMA; fp32 accumulation in TMEM) ---- */

/* The prologue of dprnn_gemm_f32 as a pass of its own (the tensor-core GEMM feeds its operands from shared memory
 * straight into the MMA): out[row,c] = (a*p_scale[b,c] + p_shift[b,c]) * rowscale[row] + p_add[b,c], b = row/rows_per_utt
 * - bottleneck norm + speaker fusion, dprnn_spe.py:136-143, and the speaker encoder's GroupNorm, dprnn_spe.py:116. */
int dprnn_prologue_apply(const float* a, float* out, long rows, int C, long rows_per_utt, const float* p_scale,
                         const float* p_shift, const float* p_add, const float* rowscale, void* stream);


/* fp32 -> bf16 (round to nearest even) copy of an activation tensor. */
int dprnn_cast_bf16(const float* x, void* out, long elems, void* stream);
int dprnn_cast_h16(const float* x, void* out, long elems, int h16, void* stream);

/* Pointwise contraction on tensor cores: C[M,N] (fp32) = epi(A[M,K] @ W[N,K]^T + bias), W in nn.Linear / Conv1d
 * layout (K-major, no transpose).  a_is_bf16 != 0: A and W are bf16 (the Linear after each LSTM,
 * src/models/dprnn.py:61,70); == 0: A and W are fp32 read as TF32 (conv2d / out / gate / end_conv1x1 of the mask
 * head, dprnn.py:155-160, and the speaker ResNet convolutions, dprnn_spe.py:17-18,27,121).
 * N in {64,128,256}; K*elem_size a multiple of 128 bytes.  DPRNN_EPI_GATED: N = 2F, W rows = [out; gate],
 * C[M,F] = tanh(out) * sigmoid(gate).  stats_partial (may be NULL; dprnn_gemm_tc_stats_bytes(M) bytes): the
 * epilogue also emits per-row sums so that mean_rstd [M/rows_per_utt, 2] of the FOLLOWING GroupNorm(1,N) / gLN
 * (eps) is produced without another pass over C.  With stats_partial == NULL, rows_per_utt > 0 makes bias a
 * per-utterance bias [M/rows_per_utt, N]. */
size_t dprnn_gemm_tc_stats_bytes(int M);
int dprnn_gemm_tc(const void* A, int a_is_bf16, const void* W, const float* bias, float* C, long ldc, int M, int N,
                  int K, int epilogue, void* stats_partial, long rows_per_utt, float eps, float* mean_rstd,
                  void* stream);

/* dprnn_gemm_tc as a persistent, pipelined kernel for the many-row 1x1 convolutions (W resident in shared memory,
 * dynamic tile scheduling, TMEM double buffering, staged TMA stores): C[M,N_out] = epi(A @ W^T + bias).
 * bias: [N], or per utterance [*, N] selected by row / bias_rows_per_utt (> 0) or bias_row_utt[row] (ragged);
 * DPRNN_EPI_AFFINE_PRELU: prelu(acc * post_scale[n] + post_shift[n]).  dprnn_gemm_persist_supported() tells whether
 * (operand type, N, K, epilogue) is built; workspace: dprnn_gemm_persist_workspace_bytes() bytes (scheduler ticket). */
/* a_kind: DPRNN_GEMM_TF32 - A, W fp32 read (truncated) as TF32; DPRNN_GEMM_BF16 - A, W bf16; DPRNN_GEMM_F32X2 - A fp32,
 * split in shared memory into bf16 pairs hi + lo (16 significand bits), W packed by the caller as [N, 2K] bf16 holding,
 * for every 32 consecutive k, hi(32) then lo(32) (hi = bf16(w), lo = bf16(w - hi)); three MMAs per K slice.  The tolerance
 * ('fp16') mode uses F32X2 for every 1x1 convolution: TF32 truncation was its dominant error. */
#define DPRNN_GEMM_TF32 0
#define DPRNN_GEMM_BF16 1
#define DPRNN_GEMM_F32X2 2
size_t dprnn_gemm_persist_workspace_bytes(void);
int dprnn_gemm_persist_supported(int a_kind, int N, int K, int epilogue);
int dprnn_gemm_persist(const void* A, int a_kind, const void* W, const float* bias, long bias_rows_per_utt,
                       const int* bias_row_utt, const float* post_scale, const float* post_shift, const float* prelu_a,
                       float* C, long ldc, int M, int N, int K, int epilogue, void* workspace, void* stream);

/* Both weight gradients and the bias gradient of one LSTM direction from ONE pass over its d gates (backward of
 * src/models/dprnn.py:23-28):  C1[N1,128] (+)= A^T B1,  C2[N1,128] (+)= A^T shift_t(B2),  colsum[N1] (+)= column sums of A.
 * A [rows, lda] (pointer at the direction's first column, N1 % 128 == 0 columns used), B1 [rows, ldb1] and B2 [rows, ldb2]
 * (128 columns each); rows are the chunk positions of a [B, S, K] batch, row = (b*S + s)*K + k.  inter = 0: time runs
 * along k (intra-chunk layer), 1: along s.  shift in {-1, 0, +1}: B2 is read at time t + shift and is zero outside the
 * sequence (h_{t-1} for the forward direction, h_{t+1} for the reverse one: no shifted copy of h).  shift = 0 with
 * B1 | B2 = the two halves of h gives the Linear's dW = dy^T h and db.
 * is_bf16 = 0: fp32 operands read as TF32 (leading dimensions in floats, multiples of 32); 1: all three operands bf16
 * (d gates from dprnn_lstm_bptt_tc_bf16out, x and h as the bf16 operands of the tensor-core forward; leading dimensions in
 * elements, multiples of 64).  fp32 accumulation, fixed reduction order.  workspace: dprnn_gemm_atb_dual_workspace_bytes(N1). */
int dprnn_gemm_atb_dual_supported(int is_bf16, int N1, long lda, long ldb1, long ldb2);
size_t dprnn_gemm_atb_dual_workspace_bytes(int N1);
int dprnn_gemm_atb_dual(const void* A, int is_bf16, long lda, int N1, const void* B1, long ldb1, const void* B2, long ldb2,
                        int B, int S, int K, int inter, int shift, float* C1, long ldc1, float* C2, long ldc2, float* colsum,
                        int accumulate, int accumulate_colsum, void* workspace, void* stream);

/* Deep-K contraction of the training step's backward (d x = d gates @ W_ih: src/models/dprnn.py:51-70 differentiated):
 * C[M,128] (+)= A[M,K] @ W[128,K]^T; a_is_bf16 = 0: fp32 operands read as TF32, K % 32 == 0; 1: bf16 operands (the d gates
 * dprnn_lstm_bptt_tc_bf16out writes), K % 64 == 0.  Any depth: W streams with A, 256-row tiles share every W block.
 * accumulate != 0 adds into C through TMA reduce-add (each element of C is touched once: deterministic).  lda in
 * elements, ldc in floats.  workspace: dprnn_gemm_kdeep_workspace_bytes() bytes (scheduler ticket). */
size_t dprnn_gemm_kdeep_workspace_bytes(void);
int dprnn_gemm_kdeep_supported(int a_is_bf16, int N, int K, long lda, long ldc);
int dprnn_gemm_kdeep(const void* A, int a_is_bf16, long lda, const void* W, float* C, long ldc, int M, int N, int K,
                     int accumulate, void* workspace, void* stream);

/* 1x1 conv -> BatchNorm1d (eval: per-channel scale/shift from dprnn_batchnorm_affine) -> PReLU in one pass
 * (ResBlock, src/models/dprnn_spe.py:32-34): C[M,N] = prelu(A @ W^T * scale[n] + shift[n]); fp32 (TF32) operands,
 * N in {128,256}. */
int dprnn_gemm_tc_affine_prelu(const void* A, const void* W, const float* scale, const float* shift,
                               const float* prelu_a, float* C, long ldc, int M, int N, int K, void* stream);

/* The Linear(ndir*H -> 128) after each LSTM (src/models/dprnn.py:61,70) as a persistent, pipelined tcgen05 kernel:
 * C[M,128] (fp32, contiguous) = A[M,K] (bf16) @ W[128,K]^T (bf16) + bias, K in {128,256}; W stays resident in shared
 * memory, the output goes through swizzled staging + TMA stores.  stats_partial / mean_rstd as in dprnn_gemm_tc. */
int dprnn_linear_bf16_stats(const void* A, const void* W, const float* bias, float* C, int M, int K,
                            void* stats_partial, long rows_per_utt, float eps, float* mean_rstd, void* stream);

/* Same kernel with the output rounded to bf16 (C_bf16 [M,128] bf16); the statistics are taken from the fp32 values.
 * For both: mean_rstd == NULL with stats_partial != NULL leaves the per-row {sum, sumsq} for the caller to reduce
 * (dprnn_row_stats_finalize_ragged). */
int dprnn_linear_bf16out_stats(const void* A, const void* W, const float* bias, void* C_bf16, int M, int K,
                               void* stats_partial, long rows_per_utt, float eps, float* mean_rstd, void* stream);
/* ... with A, W and the output in the 16-bit format h16 (DPRNN_H16_*). */
int dprnn_linear_h16out_stats(const void* A, const void* W, const float* bias, void* C_h16, int M, int K,
                              void* stats_partial, long rows_per_utt, float eps, float* mean_rstd, int h16,
                              void* stream);

/* Linear + norm + residual of a half-block (dprnn.py:86-92 / 96-99) as ONE launch that keeps the Linear output in L2
 * (csrc/linear_normres.cu): x_h16[M,128] (16-bit, in place) += norm_u(h[M,K] @ W[128,K]^T + bias), statistics per utterance
 * of rows_per_utt (>= 128) rows.  y_scratch [M,128] 16-bit and stats_partial (dprnn_gemm_tc_stats_bytes(M) bytes) are
 * scratch, mean_rstd [M / rows_per_utt, 2] an output, workspace dprnn_linear_normres_workspace_bytes(n_utt) bytes (zeroed
 * by the call).  discard_y is a flag word: bit 0 - y's cache lines are discarded from L2 after their single use instead of
 * being written back; bits 8.. - lead: the Linear pass never runs more than that many utterances ahead of the norm pass
 * (0 = unbounded), which is what keeps y inside the 126 MB L2.  Bit for bit the results of dprnn_linear_h16out_stats
 * followed by dprnn_norm_residual_h16res. */
size_t dprnn_linear_normres_workspace_bytes(int n_utt);
int dprnn_linear_normres_h16(const void* h, const void* W, const float* bias, void* y_scratch, void* x_h16,
                             const float* gamma, const float* beta, int M, int K, void* stats_partial, long rows_per_utt,
                             float eps, float* mean_rstd, void* workspace, int discard_y, int h16, void* stream);

/* The tail of a DPRNN half-block, dprnn.py:86-92 / 96-99, as ONE persistent tcgen05 kernel (bf16 mode):
 *   y = h[M,K] (bf16) @ W[128,K]^T (bf16) + bias;  x[M,128] (fp32, in place) += (y - mean_u) * rstd_u * gamma + beta,
 * mean_u / rstd_u (biased variance, eps) over all rows x 128 columns of utterance u = rows [row_off[u], row_off[u+1])
 * (row_off: n_utt+1 int64 on the device, row_off[n_utt] == M; utterances need not be tile-aligned or equal-length);
 * x_bf16 [M,128] receives the bf16 copy of the new x (operand of the next LSTM layer).  y never exists in memory:
 * pass 0 reduces the statistics from the accumulators, pass 1 recomputes the product and applies the norm; the second
 * read of h is served from L2.  workspace: dprnn_linear_norm_workspace_bytes(M, n_utt) bytes, 256-byte aligned. */
size_t dprnn_linear_norm_workspace_bytes(int M, int n_utt);
int dprnn_linear_norm_residual_bf16(const void* h, const void* W, const float* bias, float* x, void* x_bf16,
                                    const float* gamma, const float* beta, float eps, const long* row_off, int n_utt,
                                    long max_rows_per_utt, int M, int K, void* workspace, void* stream);

/* One whole nn.LSTM layer (input projection + recurrence, both directions), src/models/dprnn.py:23-28,35-36,
 * as a fused tcgen05 kernel: per step gates = [x_t | h_{t-1}] @ [W_ih | W_hh]^T with fp32 accumulators in TMEM,
 * W resident in the shared memory of a CTA pair (cta_group::2), x_t tiles fed by TMA, cell state in registers.
 * x [rows,128] bf16 (rows = (b,s,k) chunk positions), hout [rows, ndir*128] bf16.
 * w_packed [ndir*512, 256] bf16: for direction d, CTA rank r, instruction nh: 128 rows
 *   {[W_ih | W_hh][q*128 + 64*nh + j, :] : q in (2r, 2r+1), j < 64};  bias_perm[d][nh*256 + q*64 + j] =
 *   (b_ih + b_hh)[q*128 + 64*nh + j].  Rows / biases of the i, f, o gates (q = 0, 1, 3) are pre-scaled by 1/2 (the
 *   kernel evaluates sigmoid(x) = 1/2 tanh(x/2) + 1/2).
 * inter == 0: sequences (b,s) run along k (intra-chunk); inter == 1: sequences (b,k) run along s.
 * fast_act != 0: tanh.approx-based activations (1 MUFU op each); 0: expf/tanhf. hidden must be 128.
 * For the *_pp entry points below the same argument is a flag word: DPRNN_LSTM_FAST_ACT | DPRNN_LSTM_FP16, the latter
 * meaning that x, w_packed and hout are fp16 instead of bf16 (inference only). */
#define DPRNN_LSTM_FAST_ACT 1
#define DPRNN_LSTM_FP16 2
#define DPRNN_LSTM_HALF_TILES 4     /* *_pp: force 128-sequence pair tiles (default: chosen when they fit one wave) */
#define DPRNN_LSTM_FULL_TILES 8     /* *_pp: force 256-sequence pair tiles */
#define DPRNN_LSTM_DIRECT_SAVE 16   /* *_train_pp on 128-sequence tiles: the saved gates / c / h are stored from the registers
                                     * (as the 256-sequence tiles do) instead of through staging buffers + TMA (A/B knob) */
int dprnn_lstm_layer_bf16(const void* x, const void* w_packed, const float* bias_perm, void* hout, int B, int S,
                          int K, int inter, int hidden, int ndir, int fast_act, void* stream);

/* nn.Linear(2H -> F) after each LSTM, src/models/dprnn.py:61,70,86,96, on tensor cores:
 * C[M,N] (fp32, row stride ldc) = A[M,K] (bf16 row-major) @ W[N,K]^T (bf16, nn.Linear layout) + bias[N].
 * (N,K) in {(128,256),(128,128),(64,128),(256,128)}. */
int dprnn_linear_bf16(const void* A, const void* W, const float* bias, float* C, long ldc, int M, int N, int K,
                      void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Ragged (variable-length) batches - SURVEY.md section 8d cfg 3: the reference's test loop runs B = 1 on full-length
 * utterances (src/inferencers/inferencer_spe.py:25-45); these entry points give the same per-utterance results for
 * utterances of different lengths packed back to back (no padding enters a statistic, recurrence or softmax).
 *   frame space: utterance b = rows [frame_off[b], frame_off[b]+T_b); its first L_b = T_b-(ksz-1) rows are frames, the
 *                rest junk (the encoder runs over the packed waveform, stride 1); frame_utt[row] = b (int32).
 *   chunk space: utterance b = chunks [chunk_off[b], chunk_off[b]+S_b), K rows each; chunk_utt[chunk] = b (int32).
 * All index arrays live on the device; offsets / lengths are int64.  Row-wise stages use the uniform entry points.
 * --------------------------------------------------------------------------------------------------------- */

/* GroupNorm(1,C) / gLN statistics of rows [off[b], off[b]+len[b]) x C (dprnn.py:130-136, dprnn_spe.py:116,136). */
size_t dprnn_utt_stats_ragged_workspace_bytes(int B);
int dprnn_utt_stats_ragged(const float* x, int C, const long* off, const long* len, int B, float eps, void* workspace,
                           float* mean_rstd, void* stream);
/* mean/rstd per utterance (rows [row_off[b], row_off[b+1])) from the per-row sums of dprnn_linear_bf16*_stats. */
int dprnn_row_stats_finalize_ragged(const void* stats_partial, const long* row_off, int B, int cols, float eps,
                                    float* mean_rstd, void* stream);
/* dprnn_norm_residual on the packed chunk space; y_is_bf16: 0 = y fp32, 1 = y (and x_bf16) bf16, 2 = y (and x_bf16) fp16. */
int dprnn_norm_residual_ragged(const void* y, int y_is_bf16, float* x, const float* mean_rstd, const float* gamma,
                               const float* beta, const int* chunk_utt, long total_chunks, int K, int C, void* x_bf16,
                               void* stream);
int dprnn_norm_residual_ragged_bf16res(const void* y_bf16, void* x_bf16, float* x_f32_out, const float* mean_rstd,
                                       const float* gamma, const float* beta, const int* chunk_utt, long total_chunks,
                                       int K, int C, void* stream);     /* dprnn_norm_residual_bf16res, packed chunk space */
int dprnn_norm_residual_ragged_h16res(const void* y_h16, void* x_h16, float* x_f32_out, const float* mean_rstd,
                                      const float* gamma, const float* beta, const int* chunk_utt, long total_chunks,
                                      int K, int C, int h16, void* stream);
/* dprnn_unfold / dprnn_fold_prelu between the packed frame and chunk spaces (integer maps bit-exact per utterance). */
int dprnn_unfold_ragged(const float* y, float* x, const int* chunk_utt, const long* chunk_off, const long* frame_off,
                        const long* L, long total_chunks, int K, int P, int F, void* stream);
int dprnn_fold_prelu_ragged(const float* x, float* out, const int* frame_utt, const long* frame_off, const long* L,
                            const long* chunk_off, const long* S, long total_rows, int K, int P, int F,
                            const float* prelu_a, void* stream);
/* dprnn_unfold_ragged also writing the 16-bit copy x_h16 (the rounding of dprnn_cast_h16); x may be NULL when the residual
 * stream is kept in 16 bits only. */
int dprnn_unfold_ragged_h16(const float* y, float* x, void* x_h16, const int* chunk_utt, const long* chunk_off,
                            const long* frame_off, const long* L, long total_chunks, int K, int P, int F, int h16,
                            void* stream);
/* dprnn_norm_residual_fold_prelu_h16 on packed ragged batches: bit for bit dprnn_norm_residual_ragged_h16res (fp32 output)
 * followed by dprnn_fold_prelu_ragged, without the fp32 chunk-space tensor between them. */
int dprnn_norm_residual_fold_prelu_ragged_h16(const void* y_h16, const void* x_h16, const float* mean_rstd,
                                              const float* gamma, const float* beta, float* out, const int* frame_utt,
                                              const long* frame_off, const long* L, const long* chunk_off, const long* S,
                                              long total_rows, int K, int P, int F, const float* prelu_a, int h16,
                                              void* stream);
/* dprnn_mask_decode for stride 1: out[frame_off[b]+t], t < T_b. */
int dprnn_mask_decode_ragged(const float* mask, const float* enc, const float* wdec, float* out, const int* frame_utt,
                             const long* frame_off, const long* L, long total_rows, int N, int ksz, void* stream);
/* dprnn_att_rowscale; scores of utterance b are kept at rows frame_off[b] .. +La[b] of a frame-space scratch. */
int dprnn_att_rowscale_ragged(const float* enc, const float* s1, const float* s0, const float* wavg, const float* bavg,
                              const float* v, float* scores, float* rowscale, const int* frame_utt,
                              const long* frame_off, const long* L, const long* La, int B, long total_rows, int N,
                              int ksz, void* stream);
/* dprnn_affine_add_prelu_pool3: out row r (utterance out_utt[r]) = max over rows in_off[b]+3*(r-out_off[b])+{0,1,2}. */
int dprnn_affine_add_prelu_pool3_ragged(const float* y, const float* scale, const float* shift, const float* skip,
                                        const float* prelu_a, float* out, const int* out_utt, const long* in_off,
                                        const long* out_off, long total_out_rows, int C, void* stream);
int dprnn_time_sum_ragged(const float* x, float* emb, const long* off, const long* len, int B, int C, const float* div,
                          void* stream);
/* dprnn_gemm_f32 with the utterance of every row given explicitly (prologue / per-utterance bias). */
int dprnn_gemm_f32_ragged(const float* A, long lda, const float* Wt, long ldw, float* C, long ldc, int M, int N, int K,
                          const float* bias, int bias_per_utt, float bias_scale, const int* row_utt,
                          const float* p_scale, const float* p_shift, const float* p_add, const float* rowscale,
                          int epilogue, void* stream);
/* dprnn_prologue_apply / dprnn_gemm_tc (per-utterance bias [n_utt,N]) with the utterance of every row given explicitly. */
int dprnn_prologue_apply_ragged(const float* a, float* out, long rows, int C, const int* row_utt, const float* p_scale,
                                const float* p_shift, const float* p_add, const float* rowscale, void* stream);
int dprnn_gemm_tc_ragged(const void* A, int a_is_bf16, const void* W, const float* bias_per_utt, const int* row_utt,
                         float* C, long ldc, int M, int N, int K, int epilogue, void* stream);
/* dprnn_lstm_layer_bf16 with two half-jobs (64 rows per CTA each, cta_group::2 M = 128) per CTA pair in ping-pong, so that
 * one half-job's MMA -> epilogue hand-off runs under the other's cell update (csrc/lstm_tc_pp.cu).  Same arguments and
 * results; w_packed rows for direction d, CTA rank r, MMA nh: {[W_ih | W_hh][gate*H + 64*nh + 32*r + u] : gate < 4, u < 32}
 * (same 1/2 pre-scale of the i, f, o rows); bias_perm as dprnn_lstm_layer_bf16. */
int dprnn_lstm_layer_bf16_pp(const void* x, const void* w_packed, const float* bias_perm, void* hout, int B, int S, int K,
                             int inter, int hidden, int ndir, int fast_act, void* stream);
/* ... on the input x_in + norm(y): the norm + residual that ends the previous half-block (src/models/dprnn.py:90-92,
 * 98-99; mean_rstd [B,2] per utterance, gamma / beta [128]) is applied to every input tile in shared memory before the
 * tensor core reads it, and x_out [rows,128] (16-bit, NOT x_in) receives the updated residual stream.  Results are bit for
 * bit those of dprnn_norm_residual_h16res followed by dprnn_lstm_layer_bf16_pp; the stand-alone norm pass (2.4 GB per
 * half-block at B = 64) disappears.  Uniform batches, inference. */
int dprnn_lstm_layer_bf16_pp_fused(const void* x_in, const void* y, const float* mean_rstd, const float* gamma,
                                   const float* beta, void* x_out, const void* w_packed, const float* bias_perm, void* hout,
                                   int B, int S, int K, int inter, int hidden, int ndir, int flags, void* stream);
int dprnn_lstm_inter_bf16_ragged_pp(const void* x, const void* w_packed, const float* bias_perm, void* hout,
                                    long total_chunks, int K, const void* utt_jobs, int n_utt, int hidden, int ndir,
                                    int fast_act, void* stream);
/* dprnn_lstm_layer_bf16_pp as a PERSISTENT kernel over time-sliced jobs (csrc/lstm_tc_sliced.cu): every (tile, direction)
 * job is cut into nslices slices of ceil(T / nslices) steps which <= 74 resident CTA pairs draw by atomic ticket, turning
 * ceil(jobs / 74) waves into ceil(nslices * jobs / 74) / nslices.  Same arguments (flags: DPRNN_LSTM_*), weight packing and
 * results (bit for bit) as dprnn_lstm_layer_bf16_pp; nslices <= 0: chosen by dprnn_lstm_sliced_auto; max_pairs > 0 caps
 * the resident pairs; workspace: dprnn_lstm_sliced_workspace_bytes(...) bytes, 256-byte aligned, private to the call while
 * it runs (scheduler ticket, per-job completion counters, the cell-state hand-off scratch). */
size_t dprnn_lstm_sliced_workspace_bytes(int B, int S, int K, int inter, int ndir);
int dprnn_lstm_sliced_auto(int B, int S, int K, int inter, int ndir, int pairs, int kmax);
int dprnn_lstm_layer_bf16_sliced(const void* x, const void* w_packed, const float* bias_perm, void* hout, int B, int S, int K,
                                 int inter, int hidden, int ndir, int flags, int nslices, int max_pairs, void* workspace,
                                 void* stream);
/* Inter-chunk layer of dprnn_lstm_layer_bf16 on the packed chunk space: utt_jobs = n_utt x {int32 first chunk, int32
 * number of chunks}, in the order the pair-jobs should be scheduled (longest first). */
int dprnn_lstm_inter_bf16_ragged(const void* x, const void* w_packed, const float* bias_perm, void* hout,
                                 long total_chunks, int K, const void* utt_jobs, int n_utt, int hidden, int ndir,
                                 int fast_act, void* stream);

/* RawNet3 front-end of DPRNN-RawNet (src/models/rawnet/RawNet3.py:23-32,76-83, RawNetBasicBlock.py:8-28; cfg 4):
 * PreEmphasis -> InstanceNorm1d(1, eps 1e-4, affine in_w/in_b) -> ParamSincFB(n_filters, kernel, stride) -> abs -> log(.+1e-6)
 * -> minus time mean.  wave [B,T] (16 kHz) -> out [B, T', n_filters] channels-last, T' = (T-kernel)/stride+1.
 * low_hz/band_hz [n_filters/2], window/n_half [kernel/2]: the filterbank's parameters and buffers; filt_scratch
 * [kernel*n_filters] floats, stats_scratch [2B] floats.  The filter formula restates asteroid_filterbanks 0.4.0
 * (third-party; parity unpinned). */
int dprnn_rawnet_frontend(const float* wave, int B, long T, const float* in_w, const float* in_b, const float* low_hz,
                          const float* band_hz, const float* window, const float* n_half, int n_filters, int kernel,
                          int stride, float sample_rate, float* filt_scratch, float* stats_scratch, float* out,
                          void* stream);

/* ---- RawNet3 Res2Net blocks + attentive statistics pooling (src/models/rawnet/RawNetBasicBlock.py:111-142,
 * RawNet3.py:88-134), channels-last [B*T, C] activations ---- */

/* conv (as a contraction over K, any K with K*4 % 128 == 0) -> + bias -> ReLU -> BatchNorm-eval affine (-> + residual):
 * C[M,N] = relu(A @ W^T + bias) * scale[n] + shift[n] (+ residual[row*ldres + n]); TF32 operands, N in {128, 256};
 * scale/shift may be NULL (plain ReLU); bias_rows_per_utt > 0: bias is per utterance [M/rows, N].  (Bottle2neck: conv1/bn1, convs[i]/bns[i], conv3/bn3 + residual; layer4;
 * attention[0..2].) */
int dprnn_gemm_tc_relu_affine(const void* A, const void* W, const float* bias, long bias_rows_per_utt,
                              const float* scale, const float* shift, const float* residual, long ldres, float* C,
                              long ldc, int M, int N, int K, void* stream);
/* Exact-fp32 twin on CUDA cores (parity mode); Wt = transposed weight [K, N] with row stride ldw, A / C may be column
 * slices (lda / ldc). */
int dprnn_gemm_f32_relu_affine(const float* A, long lda, const float* Wt, long ldw, const float* bias,
                               long bias_rows_per_utt, const float* scale, const float* shift, const float* residual,
                               long ldres, float* C, long ldc, int M, int N, int K, void* stream);
/* im2col of a kernel-3 dilated conv with the Res2Net input sum: col[r, tap*C + c] = (a + b)[r + (tap-1)*dil, c] inside
 * the utterance (T rows each), 0 outside; b may be NULL.  a, b: column slices of [rows, ld*] buffers. */
int dprnn_res2_gather(const float* a, long lda, const float* b, long ldb, float* col, long rows, long T, int C, int dil,
                      void* stream);
/* MaxPool1d(k) over time, optional second operand added first: out[b,t',c] = max_i (x (+ y))[b, k*t'+i, c];
 * out may be a column slice (ldo). */
int dprnn_maxpool_time(const float* x, const float* y, float* out, long ldo, int B, long T, int C, int k, void* stream);
/* per (utterance, channel) over T rows: mean -> mean[b,c]; with std != NULL also sqrt(clamp(unbiased var, 1e-4, 1e4)). */
int dprnn_col_mean_std(const float* x, float* mean, float* std, int B, long T, int C, void* stream);
/* AFMS (RawNetBasicBlock.py:48-55): out = (x + alpha[c]) * gate[b,c]; out may be a column slice (ldo). */
int dprnn_afms_apply(const float* x, const float* alpha, const float* gate, float* out, long ldo, int B, long T, int C,
                     void* stream);
/* out = a + b elementwise (n % 4 == 0). */
int dprnn_add2(const float* a, const float* b, float* out, long n, void* stream);
/* out[b,c] = act(x[b,c] * scale[c] + shift[c]); act: 0 none, 1 sigmoid. */
int dprnn_affine_vec(const float* x, const float* scale, const float* shift, float* out, int B, int C, int act,
                     void* stream);
/* Attentive statistics pooling (RawNet3.py:119-124): w = softmax over time of logits[b,:,c];
 * out[b, c] = sum_t x w, out[b, C + c] = sqrt(clamp(sum_t x^2 w - mu^2, 1e-4, 1e4)). */
int dprnn_att_stats_pool(const float* x, const float* logits, float* out, int B, long T, int C, void* stream);

/* Polyphase FIR resampling of the reference utterance before RawNet3 (torchaudio.transforms.Resample(8000, 16000) in
 * src/inferencers/inferencer_rawnet.py:21-23,36 and src/trainers/trainer_rawnet.py:14-16,31): x [B,T] -> out [B,To],
 * out[b, q*nw + i] = sum_k kernel[i,k] * x[b, q*orig + k - width] (zero outside), kernel [nw, taps] from the host
 * (windowed-sinc, tss_with_dprnn_b200/resample.py). */
int dprnn_resample_fir(const float* x, const float* kernel, float* out, int B, long T, long To, int orig, int nw,
                       int taps, int width, void* stream);

/* ---- the callers' side of the path ---- */

/* 16-bit PCM -> float32 with the normalisation of soundfile.read(dtype='float32') (x / 32768; the reference's datasets
 * read their wav segments that way, src/datasets/librimix_spe.py:50-55): pcm [n] int16 -> out [n]. */
int dprnn_pcm16_to_f32(const void* pcm, float* out, long n, void* stream);
/* SI-SDR in dB per utterance (asteroid recipe: zero-mean, eps 1e-8; src/trainers/trainer_spe.py:39,
 * src/inferencers/inferencer_spe.py:37-43).  Utterance b = samples [off[b], off[b]+len[b]) of est / target, or
 * [b*uniform_len, +uniform_len) when off == len == NULL. */
int dprnn_si_sdr(const float* est, const float* target, const long* off, const long* len, long uniform_len, int B,
                 float* out_db, void* stream);
/* The TrainerSpe loss (src/trainers/trainer_spe.py:39-43): loss = mean_b -SI-SDR(est_b, target_b) + ce_gamma *
 * mean_b CrossEntropy(logits_b, spk_b) (one source, so asteroid's PIT wrapper is the identity).  est / target [B,T],
 * logits [B,C], spk [B] int64 (logits == NULL: SI-SDR term only, as the BSS trainer).  terms [B,2] receives the per-utterance (neg SI-SDR, ce_gamma * CE); loss3 = {total,
 * SI-SDR part, CE part}; d_est [B,T] / d_logits [B,C] receive d loss / d est and d loss / d logits. */
int dprnn_train_loss(const float* est, const float* target, long T, const float* logits, int C, const long* spk,
                     float ce_gamma, int B, float* terms, float* loss3, float* d_est, float* d_logits, void* stream);
/* Two-source permutation-invariant assignment (asteroid PITLossWrapper(pairwise_neg_sisdr, pit_from='pw_mtx'),
 * src/trainers/trainer.py:39): est, target [B,2,T] -> perm [B] (0 = identity, 1 = swapped; ties -> identity),
 * target_perm [B,2,T] = the targets in matched order, pairwise [B,2,2] (may be NULL) = the neg-SI-SDR matrix.
 * The PIT loss is then dprnn_train_loss over the 2B matched rows with logits == NULL (no cross-entropy term). */
int dprnn_pit2_assign(const float* est, const float* target, int B, long T, float* target_perm, int* perm,
                      float* pairwise, void* stream);
/* torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.Adam(lr, (beta1, beta2), eps, weight_decay)
 * .step() over one flat fp32 parameter / gradient buffer (src/trainers/trainer.py:42-43,115-116; step >= 1 is Adam's
 * step count; max_norm <= 0 disables clipping).  total_norm_out[0] receives the global gradient norm. */
size_t dprnn_clip_adam_workspace_bytes(void);
int dprnn_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float max_norm, int step,
                         void* workspace, float* total_norm_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Backward of the path (cfg 5: DPRNN-Spe training step; src/trainers/trainer_spe.py:37-56), exact fp32.
 * Data gradients of the contractions reuse dprnn_gemm_f32 (dX = dY @ W with W [N_out, K_in] as the [K, N] operand);
 * dprnn_unfold / dprnn_fold_prelu(prelu_a = NULL) are each other's adjoints.
 * --------------------------------------------------------------------------------------------------------- */

/* Training forward of the recurrence: dprnn_lstm_recurrence_f32 that also stores the gate activations i,f,g,o
 * (gates [rows, ndir*4H]) and the cell state (cstate [rows, ndir*H]) of every step. */
int dprnn_lstm_recurrence_f32_train(const float* gx, const float* whhT, float* hout, float* gates, float* cstate,
                                    long nseq, int T, long seq_div, long seq_outer_stride, long seq_inner_stride,
                                    long step_stride, int hidden, int ndir, void* stream);
/* Training forward on the tensor cores: dprnn_lstm_layer_bf16 (bf16 operands [x_t | h_{t-1}] [W_ih | W_hh]^T, fp32
 * accumulation, fp32 cell state) whose epilogue also stores what BPTT needs: the gate ACTIVATIONS as bf16, packed per
 * 8-unit chunk - gates_packed [rows][ndir][16 chunks][4 gates i,f,g,o][8 units] (rows*ndir*512 bf16; element
 * (row, d, gate q, unit u) at ((row*ndir + d)*16 + u/8)*32 + q*8 + u%8) -, cstate [rows, ndir*H] and h in fp32
 * (hout_f32 [rows, ndir*H]); hout_bf16 as in dprnn_lstm_layer_bf16. */
int dprnn_lstm_layer_bf16_train(const void* x, const void* w_packed, const float* bias_perm, void* hout_bf16, void* gates_packed,
                                float* cstate, float* hout_f32, int B, int S, int K, int inter, int hidden, int ndir,
                                int fast_act, void* stream);
int dprnn_lstm_layer_bf16_train_pp(const void* x, const void* w_packed, const float* bias_perm, void* hout_bf16,
                                   void* gates_packed, float* cstate, float* hout_f32, int B, int S, int K, int inter,
                                   int hidden, int ndir, int fast_act, void* stream);   /* half-job kernel + packing */
/* BPTT: dh_out [rows, ndir*H] = gradient of the layer output; whh [ndir][4H][H] (PyTorch layout);
 * dgates [rows, ndir*4H] = gradient of the gate pre-activations of every step (same sequence geometry as forward). */
int dprnn_lstm_bptt_f32(const float* dh_out, const float* gates, const float* cstate, const float* whh, float* dgates,
                        long nseq, int T, long seq_div, long seq_outer_stride, long seq_inner_stride, long step_stride,
                        int hidden, int ndir, void* stream);
/* dprnn_lstm_bptt_f32 with the recurrent contraction d h_{t-1} = d gates_t @ W_hh on the tensor cores (tcgen05, CTA pair,
 * bf16 operands, fp32 accumulation in TMEM; everything element-wise and the d gates output stay fp32).
 * gates_packed: the bf16 layout dprnn_lstm_layer_bf16_train writes.  whhT_bf16: [ndir][H][4H] bf16 = W_hh^T per
 * direction.  fast_act: bit DPRNN_LSTM_FAST_ACT = tanh.approx for tanh(c_t), as the forward kernel; tiles of 128 sequences
 * per CTA pair (64 rows per CTA, cta_group::2 M = 128) are chosen when they still fit the SMs in one wave (small
 * batches), DPRNN_LSTM_HALF_TILES / DPRNN_LSTM_FULL_TILES force that choice. */
int dprnn_lstm_bptt_tc(const float* dh_out, const void* gates_packed, const float* cstate, const void* whhT_bf16, float* dgates,
                       long nseq, int T, long seq_div, long seq_outer_stride, long seq_inner_stride, long step_stride,
                       int hidden, int ndir, int fast_act, void* stream);
/* The same with d gates written as bf16 [rows, ndir*4H] (TMA stores of the operand tile the tensor core has just read):
 * the input format of dprnn_gemm_kdeep (a_is_bf16 = 1) and dprnn_gemm_atb_dual (is_bf16 = 1); nseq % seq_div == 0. */
int dprnn_lstm_bptt_tc_bf16out(const float* dh_out, const void* gates_packed, const float* cstate, const void* whhT_bf16,
                               void* dgates_bf16, long nseq, int T, long seq_div, long seq_outer_stride,
                               long seq_inner_stride, long step_stride, int hidden, int ndir, int fast_act, void* stream);
/* h_prev for the W_hh gradient: out[row(n,t)] = h[row(n, previous step of the direction)], 0 at the first step. */
int dprnn_shift_rows(const float* h, float* out, long nseq, int T, long seq_div, long seq_outer_stride,
                     long seq_inner_stride, long step_stride, int hidden, int ndir, void* stream);
/* C[N1,N2] (ldc) (+)= A[M,N1]^T @ B[M,N2]: weight gradients (two-stage deterministic reduction over the rows). */
size_t dprnn_gemm_atb_workspace_bytes(long M, int N1, int N2);
int dprnn_gemm_atb(const float* A, long lda, const float* B, long ldb, float* C, long ldc, long M, int N1, int N2,
                   int accumulate, void* workspace, void* stream);
/* dprnn_gemm_atb on the tensor cores: both operands are read straight from fp32 memory as MN-major TF32 tiles (TMA,
 * 128B swizzle), fp32 accumulation in TMEM, deterministic split-row reduction.  Built for one of N1 / N2 == 128 and the
 * other a multiple of 128 (the LSTM / Linear weight gradients); dprnn_gemm_atb_tc_supported() tells. */
int dprnn_gemm_atb_tc_supported(int N1, int N2, long lda, long ldb);
size_t dprnn_gemm_atb_tc_workspace_bytes(int N1, int N2);
int dprnn_gemm_atb_tc(const float* A, long lda, const float* B, long ldb, float* C, long ldc, long M, int N1, int N2,
                      int accumulate, void* workspace, void* stream);
/* dprnn_gemm_atb_tc for N2 = 128 with the column sums of A in the same pass: C[N1,128] (+)= A^T B, colsum[N1] (+)= sum_m
 * A[m,:] (dW_ih = dgates^T x with db = sum dgates): a 32-column group of ones extends B in shared memory, so the bias
 * gradient costs no pass of its own over A.  workspace: dprnn_gemm_atb_tc_colsum_workspace_bytes(N1) bytes. */
int dprnn_gemm_atb_tc_colsum_supported(int N1, int N2, long lda, long ldb);
size_t dprnn_gemm_atb_tc_colsum_workspace_bytes(int N1);
int dprnn_gemm_atb_tc_colsum(const float* A, long lda, const float* B, long ldb, float* C, long ldc, float* colsum, long M,
                             int N1, int N2, int accumulate, int accumulate_colsum, void* workspace, void* stream);
/* out[n] (+)= sum_m X[m,n] (* Y[m,n] if Y): bias gradients, BatchNorm reductions. */
size_t dprnn_col_sum_workspace_bytes(int N);
int dprnn_col_sum(const float* X, long ldx, const float* Y, long ldy, long M, int N, float* out, int accumulate,
                  void* workspace, void* stream);
/* GroupNorm(1,C) / gLN backward per utterance: z = gamma*(y-mean)*rstd + beta; given dz: dy (=, or += if
 * accumulate_dy), dgamma += , dbeta += . */
size_t dprnn_gn_bwd_workspace_bytes(int B, int C);
int dprnn_groupnorm_bwd(const float* dz, const float* y, const float* mean_rstd, const float* gamma, int B,
                        long rows_per_utt, int C, float* dy, int accumulate_dy, float* dgamma, float* dbeta,
                        void* workspace, void* stream);
/* The same, also writing a bf16 copy of dy (the operand of the Linear's bf16 weight-gradient pass and of d h = dy W);
 * dy may then be NULL (accumulate_dy = 0): only the bf16 copy is written. */
int dprnn_groupnorm_bwd_h16(const float* dz, const float* y, const float* mean_rstd, const float* gamma, int B,
                            long rows_per_utt, int C, float* dy, int accumulate_dy, float* dgamma, float* dbeta,
                            void* workspace, void* dy_bf16, void* stream);
/* PReLU adjoint: dx = dy * (x > 0 ? 1 : a); da[0] += sum dy*x*[x <= 0]; workspace: 148*16 doubles. */
int dprnn_prelu_bwd(const float* dy, const float* x, const float* prelu_a, float* dx, long n, float* da, void* workspace,
                    void* stream);
/* MaxPool1d(3) adjoint: dv [B,Lin,C] from dy [B,Lin/3,C] and the pooled input v (first maximum wins). */
int dprnn_pool3_bwd(const float* dy, const float* v, float* dv, int B, long Lin, int C, void* stream);
/* tanh(po)*sigmoid(pg) adjoint; pre = [po | pg] [rows, 2F] -> dpre [rows, 2F]. */
int dprnn_gated_bwd(const float* dg, const float* pre, float* dpre, long rows, int F, void* stream);
/* its forward with unpacked halves: out [rows,F] = tanh(pre[:, :F]) * sigmoid(pre[:, F:]) (training forward keeps pre). */
int dprnn_gated_fwd(const float* pre, float* out, long rows, int F, void* stream);
int dprnn_mul(const float* a, const float* b, float* out, long n, void* stream);
int dprnn_axpy(const float* a, float alpha, float* out, long n, int accumulate, void* stream);
/* dpre = dy * act'(y) from the activation OUTPUT y: act 1 = ReLU, 2 = sigmoid. */
int dprnn_act_bwd(const float* dy, const float* y, float* dpre, long n, int act, void* stream);
/* Decoder (ConvTranspose1d N->1, kernel 2, stride 1) adjoint wrt its input: dz[b,l,c] = dest[b,l] w[c,0] + dest[b,l+1] w[c,1]. */
int dprnn_decoder_bwd(const float* dest, const float* wdec, float* dz, int B, long L, int N, void* stream);
/* dw[c,j] (+)= sum_{b,l} z[b,l,c] * sig[b,l+j]: weight gradient of the kernel-2 stride-1 decoder / encoder. */
size_t dprnn_convw2_workspace_bytes(int N);
int dprnn_convw2_grad(const float* z, const float* sig, int B, long L, int N, float* dw, int accumulate, void* workspace,
                      void* stream);
/* out[b,c] = sum_l X[b,l,c] (* Y[b,l,c]); out[b,l,c] (+)= v[b,c] * (X ? X[b,l,c] : 1). */
size_t dprnn_utt_col_sum_workspace_bytes(int B, int C);
int dprnn_utt_col_sum(const float* X, const float* Y, int B, long L, int C, float* out, void* workspace, void* stream);
/* Attention-fusion backward (src/models/dprnn_spe.py:177-183,217-225; fused = n * v[b,c] * r[b,l], r = 1 + softmax(s)[src(l)],
 * s[b,j] = sum_c v * avg(n)):  dprnn_row_dot3: out[row] = sum_c A*Bm*v[utt] (= d r);  dprnn_att_softmax_bwd: upsample
 * adjoint + softmax adjoint per utterance, w2 [B,L] = d s spread back over the frames of the average conv (ds [B,La]
 * scratch/out);  dprnn_att_bwd_apply: g = dfused*r + w2*wavg[c, l%ksz], dn = v*g (gradient of the normalised encoding),
 * tdv = n*g (its sum over time is the gradient of v = fusion_linear(e)). */
int dprnn_row_dot3(const float* A, const float* Bm, const float* v, long rows, long rows_per_utt, int C, float* out,
                   void* stream);
int dprnn_att_softmax_bwd(const float* dr, const float* a, int B, long L, int ksz, float* w2, float* ds, void* stream);
int dprnn_att_bwd_apply(const float* dfused, const float* n, const float* v, const float* r, const float* w2,
                        const float* wavg, int B, long L, int C, int ksz, float* dn, float* tdv, void* stream);
int dprnn_bcast_mul(const float* v, const float* X, float* out, int B, long L, int C, int accumulate, void* stream);
/* BatchNorm1d (train) adjoint: dy = gamma*rstd*(dout - m1 - yhat*m2), m1 = mean(dout), m2 = mean(dout*yhat) per channel. */
int dprnn_bn_bwd_apply(const float* dout, const float* y, const float* mean, const float* rstd, const float* gamma,
                       const float* m1, const float* m2, float* dy, long rows, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPRNN_B200_H */
