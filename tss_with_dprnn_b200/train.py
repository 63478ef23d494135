"""Training step of the separation path (cfg 5: DPRNN-Spe, src/trainers/trainer_spe.py:14-72; DPRNN-TasNet:
src/trainers/trainer.py:100-118): forward in train mode (BatchNorm batch statistics, running-stat updates) that keeps
what the backward needs, and the hand-written backward (csrc/backward.cu, csrc/lstm_simt.cu, csrc/lstm_bptt_tc.cu,
csrc/atb_tc.cu).  precision 'fp32': exact fp32 on CUDA cores; 'bf16': LSTM forward / BPTT contractions on tcgen05 with
bf16 operands, Linear / dX / weight-gradient contractions in TF32 (cfg 5: "bf16 gate GEMMs").  The loss lives in the
caller (asteroid's PIT / SI-SDR wrapper + CrossEntropyLoss in the reference trainer), so the model is exposed to autograd
as ONE ``torch.autograd.Function``: forward -> (est, logits), backward(d_est, d_logits) -> parameter gradients;
``SpeTrainStep`` runs the whole iteration (loss kernel, all-reduce, fused clip + Adam) as device work.

Memory: per half-block the forward keeps the LSTM gates / cell state / output, the Linear output and the half-block's
input (14 A, A = one [B,S,K,128] fp32 tensor = 0.40 GB at B = 16).  With DPRNN_TRAIN_KEEP_INPUTS=0 the inputs are not kept:
the residual stream is reversible (x_in = x_out - norm(y)), and the backward walks it back while it walks the blocks in
reverse, at the price of one more pass over it per half-block.  Either way a forward can be differentiated only once
(the saved activations are freed as the backward consumes them).

Supported: DPRNNTasNet and DPRNNSpeTasNet with fusion_type in {film, add, mul, cat, att}, 'ln' / 'gLN' norms, sigmoid /
relu mask activation, uni- or bidirectional inter-RNN, kernel_size 2 / stride 1, feature_size = hidden_size = 128; frozen
parameters are skipped.  DPRNN-Spe-IRA (the refinement iterations share the core's weights: their gradients accumulate)
and an external speaker embedding (EmbTrainFunction: DPRNN-RawNet, whose RawNet3 encoder differentiates through library
ops in rawnet.py) use the same pieces.
"""
from __future__ import annotations

import os

import torch

from ._lib import lib

EPI_NONE, EPI_RELU, EPI_SIGMOID, EPI_GATED = 0, 1, 2, 3


# experiments: 4 / 8 force 64 / 128 rows per CTA in the BPTT kernel (include/dprnn_b200.h: DPRNN_LSTM_HALF_TILES / _FULL_TILES)
_BPTT_FLAGS = int(os.environ.get('DPRNN_BPTT_FLAGS', '0')) & 12
# 4 / 8 likewise for the training forward, 16 = DPRNN_LSTM_DIRECT_SAVE (saved activations stored from the registers)
# 1 (default): the forward keeps every half-block's input (0.4 GB each at 16 x 3 s; 12 of them) and the backward reads it;
# 0: the forward updates the residual stream in place and the backward recomputes x_in = x_out - norm(y) (one more pass)
_KEEP_INPUTS = os.environ.get('DPRNN_TRAIN_KEEP_INPUTS', '1') != '0'
# 1 (default): in the bf16-d-gates mode the LSTM output is kept as bf16 only (no fp32 copy)
_NO_HF = os.environ.get('DPRNN_TRAIN_NO_HF', '1') != '0'
_SKIP_SIDE = os.environ.get('DPRNN_TRAIN_SKIP_SIDE', '0') == '1'
_FWD_FLAGS = int(os.environ.get('DPRNN_TRAIN_LSTM_FLAGS', '0')) & 28


def _st():
    return torch.cuda.current_stream().cuda_stream


class _Ops:
    """Thin wrappers that allocate outputs / workspaces (PyTorch = device memory only)."""

    def __init__(self, dev, tf32=False):
        self.dev = dev
        self.L = lib()
        self.tf32 = tf32                # big contractions on the tensor cores (TF32 operands, fp32 accumulation)
        self.persist = os.environ.get('DPRNN_TRAIN_PERSIST', '1') != '0'     # resident-weight GEMM where it applies
        self._gp_ws = {}                # its ticket word, one per stream (launches of two streams may overlap)
        self.dual = os.environ.get('DPRNN_TRAIN_DUAL', '1') != '0'           # dW_ih, dW_hh, db from one pass over d gates
        self.kdeep = os.environ.get('DPRNN_TRAIN_KDEEP', '1') != '0'         # d x = d gates @ W_ih accumulated in place
        # d gates in bf16 (what the BPTT's tensor-core tile holds anyway): d x and the weight gradients read bf16 operands
        self.dg16 = tf32 and self.dual and self.kdeep and os.environ.get('DPRNN_TRAIN_DG16', '1') != '0'

    def empty(self, *shape):
        return torch.empty(shape, device=self.dev, dtype=torch.float32)

    def gemm(self, A, Wt, M, N, K, bias=None, epi=EPI_NONE, lda=None, ldw=None, **pro):
        """C[M,N] = epi(pro(A)[M,K] @ Wt[K,N] + bias)"""
        out = self.empty(M, N // 2 if epi == EPI_GATED else N)
        self.L.call('dprnn_gemm_f32', A, lda or K, Wt, ldw or N, out, out.shape[1], M, N, K, bias,
                    int(pro.get('bias_per_utt', False)), 1.0, int(pro.get('rows_per_utt', 0)), pro.get('p_scale'),
                    pro.get('p_shift'), pro.get('p_add'), None, epi, _st())
        return out

    def mm(self, A, W, M, N, K, bias=None, epi=EPI_NONE, rows_per_utt=0, x2=False):
        """C[M,N] = epi(A[M,K] @ W[N,K]^T + bias) with W in nn.Linear layout; tensor cores in tf32 mode.
        rows_per_utt > 0: bias is per utterance, [M / rows_per_utt, N].  x2: fp32 operands as bf16 PAIRS (DPRNN_GEMM_F32X2,
        16 significand bits, three MMAs per K slice) instead of truncated TF32 - for the speaker encoder, whose scalar PReLU
        gradients are sums with heavy cancellation that the coherent TF32 truncation error does not survive."""
        if not (self.tf32 and K % 32 == 0 and N % 64 == 0 and M >= 128):
            return self.gemm(A, W.t().contiguous(), M, N, K, bias=bias, epi=epi, bias_per_utt=rows_per_utt > 0,
                             rows_per_utt=rows_per_utt)
        out = self.empty(M, N)
        step = 256 if N % 256 == 0 else (128 if N % 128 == 0 else 64)
        if step == 256 and K > 128:
            step = 128                                   # the persistent kernel keeps [step, K] of W resident
        kind = 2 if x2 and self.persist and self.L.query('dprnn_gemm_persist_supported', 2, step, K, epi) else 0
        if x2 and not kind:
            return self.gemm(A, W.t().contiguous(), M, N, K, bias=bias, epi=epi, bias_per_utt=rows_per_utt > 0,
                             rows_per_utt=rows_per_utt)
        for n0 in range(0, N, step):
            if bias is None:
                bn = None
            elif rows_per_utt:                           # [utterances, N] -> this column block's [utterances, step]
                bn = bias if N == step else bias[:, n0:n0 + step].contiguous()
            else:
                bn = bias[n0:n0 + step]
            Wn = W[n0:n0 + step]
            if kind == 2:                                # per 32 consecutive k: hi(32) then lo(32), as bf16
                w3 = Wn.detach().float().reshape(step, K // 32, 32)
                hi = w3.to(torch.bfloat16)
                Wn = torch.cat([hi, (w3 - hi.float()).to(torch.bfloat16)], -1).reshape(step, 2 * K).contiguous()
            if self.persist and self.L.query('dprnn_gemm_persist_supported', kind, step, K, epi):
                # one CTA per SM with the weight resident and a TMA ring over the rows (csrc/gemm_persist.cu)
                ws = self._gp_ws.get(_st())
                if ws is None:
                    ws = self._gp_ws[_st()] = torch.empty(self.L.query('dprnn_gemm_persist_workspace_bytes'), device=self.dev,
                                                          dtype=torch.uint8)
                self.L.call('dprnn_gemm_persist', A, kind, Wn, bn, int(rows_per_utt), None, None, None, None,
                            out.data_ptr() + 4 * n0, N, M, step, K, epi, ws, _st())
            else:
                self.L.call('dprnn_gemm_tc', A, 0, Wn, bn, out.data_ptr() + 4 * n0, N, M, step, K, epi,
                            None, int(rows_per_utt), 0.0, None, _st())
        return out

    def mm_acc(self, A, W, M, N, K, out):
        """out[M,N] += A[M,K] @ W[N,K]^T in one pass (deep-K kernel, csrc/gemm_kdeep.cu); A and W fp32 (TF32) or both bf16;
        False when it does not apply."""
        bf = int(A.dtype == torch.bfloat16)
        if not (self.tf32 and self.kdeep and M >= 256 and self.L.query('dprnn_gemm_kdeep_supported', bf, N, K, K, N)):
            return False
        ws = self._gp_ws.get(('kd', _st()))
        if ws is None:
            ws = self._gp_ws[('kd', _st())] = torch.empty(self.L.query('dprnn_gemm_kdeep_workspace_bytes'), device=self.dev,
                                                          dtype=torch.uint8)
        self.L.call('dprnn_gemm_kdeep', A, bf, K, W, out, N, M, N, K, 1, ws, _st())
        return True

    def atb(self, A, B, M, N1, N2, out, lda=None, ldb=None, ldc=None, accumulate=True):
        """out[N1,N2] (+)= A[M,N1]^T B[M,N2]"""
        if (self.tf32 and M >= 4096 and N1 == 256 and N2 == 256 and isinstance(A, torch.Tensor) and isinstance(out, torch.Tensor)
                and lda is None and ldc is None):
            # neither operand is the 128-column one the tensor-core kernel wants: two row halves of `out`
            for h in (0, 1):
                self.atb(A.data_ptr() + 4 * 128 * h, B, M, 128, N2, out.data_ptr() + 4 * 128 * h * N2, lda=N1, ldb=ldb,
                         ldc=N2, accumulate=accumulate)
            return
        if self.tf32 and M >= 4096 and self.L.query('dprnn_gemm_atb_tc_supported', N1, N2, lda or N1, ldb or N2):
            ws = torch.empty(self.L.query('dprnn_gemm_atb_tc_workspace_bytes', N1, N2), device=self.dev, dtype=torch.uint8)
            self.L.call('dprnn_gemm_atb_tc', A, lda or N1, B, ldb or N2, out, ldc or N2, M, N1, N2, int(accumulate), ws, _st())
            return
        ws = torch.empty(self.L.query('dprnn_gemm_atb_workspace_bytes', M, N1, N2), device=self.dev, dtype=torch.uint8)
        self.L.call('dprnn_gemm_atb', A, lda or N1, B, ldb or N2, out, ldc or N2, M, N1, N2, int(accumulate), ws, _st())

    def atb_colsum(self, A, B, M, N1, N2, out, colsum, lda=None, ldb=None):
        """out[N1,N2] += A^T B and colsum[N1] = column sums of A in ONE pass over A (tensor-core mode, N2 = 128); returns
        False when that kernel does not apply (the caller then runs atb + colsum)."""
        if not (self.tf32 and M >= 4096 and self.L.query('dprnn_gemm_atb_tc_colsum_supported', N1, N2, lda or N1, ldb or N2)):
            return False
        ws = torch.empty(self.L.query('dprnn_gemm_atb_tc_colsum_workspace_bytes', N1), device=self.dev, dtype=torch.uint8)
        self.L.call('dprnn_gemm_atb_tc_colsum', A, lda or N1, B, ldb or N2, out, N2, colsum, M, N1, N2, 1, 0, ws, _st())
        return True

    def atb_dual(self, A, lda, N1, B1, ldb1, B2, ldb2, B, S, K, inter, shift, C1, ldc1, C2, ldc2, colsum, bf16=False):
        """C1[N1,128] += A^T B1, C2[N1,128] += A^T shift_t(B2) and colsum[N1] = column sums of A in ONE pass over A
        (csrc/atb_tc.cu: atb_dual_kernel; rows addressed as (time, sequence), so the time shift costs no copy); operands
        fp32 (TF32) or all bf16; False when the kernel does not apply."""
        if not (self.tf32 and self.dual and B * S * K >= 4096
                and self.L.query('dprnn_gemm_atb_dual_supported', int(bf16), N1, lda, ldb1, ldb2)):
            return False
        ws = torch.empty(self.L.query('dprnn_gemm_atb_dual_workspace_bytes', N1), device=self.dev, dtype=torch.uint8)
        self.L.call('dprnn_gemm_atb_dual', A, int(bf16), lda, N1, B1, ldb1, B2, ldb2, B, S, K, int(inter), int(shift), C1, ldc1,
                    C2, ldc2, colsum, 1, 0, ws, _st())
        return True

    def colsum(self, X, M, N, out, Y=None, ldx=None, accumulate=True):
        ws = torch.empty(self.L.query('dprnn_col_sum_workspace_bytes', N), device=self.dev, dtype=torch.uint8)
        self.L.call('dprnn_col_sum', X, ldx or N, Y, ldx or N, M, N, out, int(accumulate), ws, _st())

    def utt_colsum(self, X, Y, B, L, C):
        """out[b,c] = sum_l X[b,l,c] (* Y[b,l,c])"""
        out = self.empty(B, C)
        ws = torch.empty(self.L.query('dprnn_utt_col_sum_workspace_bytes', B, C), device=self.dev, dtype=torch.uint8)
        self.L.call('dprnn_utt_col_sum', X, Y, B, L, C, out, ws, _st())
        return out

    def utt_stats(self, x, B, elems, eps):
        ws = torch.empty(self.L.query('dprnn_utt_stats_workspace_bytes', B), device=self.dev, dtype=torch.uint8)
        mr = self.empty(B, 2)
        self.L.call('dprnn_utt_stats', x, B, elems, float(eps), ws, mr, _st())
        return mr

    def gn_bwd(self, dz, y, mr, gamma, B, rows_per_utt, C, dgamma, dbeta, dy=None, accumulate_dy=False, dy16=None, only16=False):
        ws = torch.empty(self.L.query('dprnn_gn_bwd_workspace_bytes', B, C), device=self.dev, dtype=torch.uint8)
        if dy is None and not only16:
            dy = self.empty(B * rows_per_utt, C)
        if dy16 is not None:       # also (only16: only) a bf16 copy of dy
            self.L.call('dprnn_groupnorm_bwd_h16', dz, y, mr, gamma, B, rows_per_utt, C, dy, int(accumulate_dy), dgamma, dbeta,
                        ws, dy16, _st())
        else:
            self.L.call('dprnn_groupnorm_bwd', dz, y, mr, gamma, B, rows_per_utt, C, dy, int(accumulate_dy), dgamma, dbeta, ws,
                        _st())
        return dy

    def mm16(self, A16, W, M, N, K, bias=None):
        """C[M,N] (fp32) = A16[M,K] @ bf16(W[N,K])^T + bias with bf16 operands on the persistent tensor-core kernel."""
        out = self.empty(M, N)
        ws = self._gp_ws.get(_st())
        if ws is None:
            ws = self._gp_ws[_st()] = torch.empty(self.L.query('dprnn_gemm_persist_workspace_bytes'), device=self.dev,
                                                  dtype=torch.uint8)
        self.L.call('dprnn_gemm_persist', A16, 1, W.detach().to(torch.bfloat16).contiguous(), bias, 0, None, None, None, None,
                    out, N, M, N, K, EPI_NONE, ws, _st())
        return out

    def prelu_bwd(self, dy, x, a, da):
        dx = torch.empty_like(x)
        ws = torch.empty(148 * 16 * 8, device=self.dev, dtype=torch.uint8)
        self.L.call('dprnn_prelu_bwd', dy, x, a, dx, x.numel(), da, ws, _st())
        return dx

    def axpy(self, a, out, alpha=1.0, accumulate=True):
        self.L.call('dprnn_axpy', a, float(alpha), out, a.numel(), int(accumulate), _st())

    def mul(self, a, b):
        out = torch.empty_like(a)
        self.L.call('dprnn_mul', a, b, out, a.numel(), _st())
        return out


_SIDE = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=key)
    return _SIDE[key]


def _norm_params(mod):
    if hasattr(mod, 'gamma'):
        return mod.gamma, mod.beta, 1e-8
    return mod.weight, mod.bias, mod.eps


def _check_supported(model):
    cfg = model.cfg
    if model.precision == 'fp16':
        raise NotImplementedError("the training step runs in precision 'fp32' or 'bf16' (cfg 5: bf16 gate GEMMs); "
                                  "'fp16' is an inference mode")
    if cfg['kind'] not in ('spe', 'bss', 'ira', 'rawnet') or \
            (cfg['kind'] != 'bss' and cfg['fusion_type'] not in ('film', 'add', 'mul', 'cat', 'att')):
        raise NotImplementedError("training is built for DPRNNTasNet (scripts/train/config_bss.yaml), DPRNNSpeTasNet / "
                                  "DPRNNSpeIRATasNet / DPRNNRawNetTasNet with fusion_type in {'film','add','mul','cat','att'}")
    if cfg['kernel_size'] != 2 or cfg['stride'] != 1 or cfg['feature_size'] != 128 or cfg['hidden_size'] != 128 \
            or cfg['chunk_length'] != 2 * cfg['hop_length'] or cfg['input_size'] not in (32, 64, 128):
        raise NotImplementedError('training is built for the shipped geometry (kernel 2, stride 1, F = H = 128, hop = K/2)')


# --------------------------------------------------------------------------------------------------------------
# forward (train mode), in the pieces the model variants are composed of
# --------------------------------------------------------------------------------------------------------------
def _spk_fwd(model, ops, feats, B, Lr, div):
    """Speaker encoder + time mean (dprnn_spe.py:115-122,156-163), BatchNorm in train mode: feats [B,Lr,N] -> emb [B,E] and
    what its backward needs.  Its contractions stay exact fp32 in both modes: the train-mode BatchNorm chain and the scalar
    PReLU slope gradients are cancellation-heavy (TF32 there moved one slope gradient by 40 %) and the branch is < 3 % of
    the step."""
    L_, sep, st, dev = lib(), model.separation, _st(), feats.device
    N = model.cfg['input_size']
    se = sep.spk_encoder
    mr_s = ops.utt_stats(feats, B, Lr * N, se[0].eps)
    s1 = ops.empty(B, N); s0 = ops.empty(B, N)
    L_.call('dprnn_norm_affine', mr_s, se[0].weight.detach(), se[0].bias.detach(), None, s1, s0, B, N, st)
    gnf = torch.empty_like(feats)                                  # GroupNorm(feats): kept for dW of conv0
    L_.call('dprnn_prologue_apply', feats, gnf, B * Lr, N, Lr, s1, s0, None, None, st)
    O = se[1].weight.shape[0]
    x = ops.mm(gnf, se[1].weight.detach().reshape(O, N), B * Lr, O, N, bias=se[1].bias.detach(), x2=True)
    res, Lx = [], Lr
    for rb in (se[2], se[3], se[4]):
        Cin, Cout = rb.conv1.weight.shape[1], rb.conv1.weight.shape[0]
        rows = B * Lx
        rc = dict(x=x, Lx=Lx, Cin=Cin, Cout=Cout)
        ws = torch.empty(L_.query('dprnn_bn_workspace_bytes', Cout), device=dev, dtype=torch.uint8)

        def bn(y, bnm):
            scale, shift = ops.empty(Cout), ops.empty(Cout)
            L_.call('dprnn_batchnorm_affine', y, rows, Cout, bnm.weight.detach(), bnm.bias.detach(), bnm.running_mean,
                    bnm.running_var, 1, float(bnm.eps), float(bnm.momentum if bnm.momentum is not None else 0.1), ws,
                    scale, shift, st)
            bnm.num_batches_tracked += 1
            return scale, shift

        y1 = ops.mm(x, rb.conv1.weight.detach().reshape(Cout, Cin), rows, Cout, Cin, x2=True)
        sc1, sh1 = bn(y1, rb.batch_norm1)
        a1 = ops.empty(rows, Cout)
        L_.call('dprnn_affine_prelu', y1, sc1, sh1, rb.prelu1.weight.detach(), a1, rows, Cout, st)
        y2 = ops.mm(a1, rb.conv2.weight.detach().reshape(Cout, Cout), rows, Cout, Cout, x2=True)
        sc2, sh2 = bn(y2, rb.batch_norm2)
        if hasattr(rb, 'conv_downsample'):
            skip = ops.mm(x, rb.conv_downsample.weight.detach().reshape(Cout, Cin), rows, Cout, Cin, x2=True)
        else:
            skip = x
        Lo = Lx // 3
        out = ops.empty(B, Lo, Cout)
        L_.call('dprnn_affine_add_prelu_pool3', y2, sc2, sh2, skip, rb.prelu2.weight.detach(), out, B, Lx, Cout, st)
        rc.update(y1=y1, sc1=sc1, sh1=sh1, a1=a1, y2=y2, sc2=sc2, sh2=sh2, skip=skip)
        res.append(rc)
        x, Lx = out.view(B * Lo, Cout), Lo
    E = se[5].weight.shape[0]
    C5 = se[5].weight.shape[1]
    z5 = ops.mm(x, se[5].weight.detach().reshape(E, C5), B * Lx, E, C5, bias=se[5].bias.detach(), x2=True)
    emb = ops.empty(B, E)
    L_.call('dprnn_time_sum', z5, emb, B, Lx, E, div, st)
    return emb, dict(feats=feats, mr_s=mr_s, gnf=gnf, res=res, x3=x, L3=Lx, Lr=Lr, div=div, B=B)


def _emb_linear(ops, emb, m, w_off=0, K=None):
    """m(emb[:, :K]) with the weight columns [w_off, w_off + K) of the nn.Linear m."""
    B, E = emb.shape
    Kin = m.weight.shape[1]
    K = E if K is None else K
    out = ops.empty(B, m.weight.shape[0])
    lib().call('dprnn_small_linear', emb, E, m.weight.detach().data_ptr() + 4 * w_off, Kin,
               m.bias.detach() if w_off == 0 else None, out, out.shape[1], B, out.shape[1], K, 0, _st())
    return out


def _core_fwd(model, ops, enc, mr_e, emb, B, Lm, spks):
    """Bottleneck norm + fusion + 1x1 conv, segmentation, the DPRNN blocks, PReLU, overlap-add, conv2d, gated head, end
    conv + activation (dprnn.py:166-187 / dprnn_spe.py:125-154,231-248): enc [B,Lm,N] (+ emb [B,E]) -> one mask [B*Lm,N]
    per requested speaker, and everything the backward needs."""
    L_, cfg, sep, st, dev = lib(), model.cfg, model.separation, _st(), enc.device
    N, F, H, K, P = cfg['input_size'], cfg['feature_size'], cfg['hidden_size'], cfg['chunk_length'], cfg['hop_length']
    gamma, beta, _ = _norm_params(sep.bottleneck[0])
    ft = cfg['fusion_type'] if emb is not None else None
    E = emb.shape[1] if emb is not None else 0
    mulc = addc = None
    bw = sep.bottleneck[1].weight.detach().reshape(F, -1)
    bias, bias_per_utt = sep.bottleneck[1].bias.detach(), False
    if ft == 'film':
        mulc, addc = _emb_linear(ops, emb, sep.fusion_linear_1), _emb_linear(ops, emb, sep.fusion_linear_2)
    elif ft == 'add':
        addc = _emb_linear(ops, emb, sep.fusion_linear)
    elif ft == 'mul':
        mulc = _emb_linear(ops, emb, sep.fusion_linear)
    elif ft == 'cat':
        bias = ops.empty(B, F)
        L_.call('dprnn_small_linear', emb, E, bw.data_ptr() + 4 * N, N + E, sep.bottleneck[1].bias.detach(), bias, F, B, F, E, 0, st)
        bias_per_utt = True
    rowscale = att_a = None
    if ft == 'att':                                                 # dprnn_spe.py:177-183,217-225
        ksz = cfg['kernel_size']
        mulc = _emb_linear(ops, emb, sep.fusion_linear)
        n1, n0 = ops.empty(B, N), ops.empty(B, N)
        L_.call('dprnn_norm_affine', mr_e, gamma.detach(), beta.detach(), None, n1, n0, B, N, st)
        att_a = ops.empty(B, (Lm - ksz) // ksz + 1)                 # softmax over the averaged frames (kept for the backward)
        rowscale = ops.empty(B, Lm)
        L_.call('dprnn_att_rowscale', enc, n1, n0, sep.average.weight.detach(), sep.average.bias.detach(), mulc, att_a,
                rowscale, B, Lm, N, ksz, st)
    s1e, s0e = ops.empty(B, N), ops.empty(B, N)
    L_.call('dprnn_norm_affine', mr_e, gamma.detach(), beta.detach(), mulc, s1e, s0e, B, N, st)
    fused = torch.empty_like(enc)                                   # fusion(GroupNorm(enc)): kept for dW of the 1x1 conv
    L_.call('dprnn_prologue_apply', enc, fused, B * Lm, N, Lm, s1e, s0e, addc, rowscale, st)
    y = ops.mm(fused, bw[:, :N].contiguous(), B * Lm, F, N, bias=bias, rows_per_utt=Lm if bias_per_utt else 0)
    S = L_.query('dprnn_num_chunks', Lm, K, P)
    xs = ops.empty(B, S, K, F)
    L_.call('dprnn_unfold', y, xs, B, Lm, K, P, F, st)
    del y
    rows = B * S * K
    c = dict(B=B, Lm=Lm, emb=emb, mulc=mulc, addc=addc, fused=fused, S=S, rows=rows, rowscale=rowscale, att_a=att_a)

    # ---- DPRNN blocks (dprnn.py:79-99)
    halves = []
    xb_next, n_half = None, 0
    # bf16 d gates in the backward: its weight-gradient pass reads the bf16 x and h this forward works on (kept), and the
    # half-block inputs need neither be kept in fp32 nor be recomputed
    dg16 = ops.dg16 and F == 128 and H == 128 and rows >= 4096
    # ... and with every layer bidirectional h is kept as bf16 only: the Linear and its weight gradient read that copy
    nohf = (dg16 and _NO_HF and model._engine.lstm_pingpong and all(blk.intra_rnn.rnn.bidirectional and blk.inter_rnn.rnn.bidirectional for blk in sep.dprnn_blocks)
            and bool(ops.L.query('dprnn_gemm_persist_supported', 1, F, 2 * H, EPI_NONE))
            and bool(ops.L.query('dprnn_gemm_persist_supported', 1, 2 * H, F, EPI_NONE)))
    xb = hb = None
    for blk in sep.dprnn_blocks:
        for which, (rnn, linm, nm) in enumerate(((blk.intra_rnn.rnn, blk.intra_linear, blk.intra_norm),
                                                 (blk.inter_rnn.rnn, blk.inter_linear, blk.inter_norm))):
            sfx = ['', '_reverse'] if rnn.bidirectional else ['']
            nd = len(sfx)
            wih = torch.cat([getattr(rnn, 'weight_ih_l0' + s).detach() for s in sfx], 0)              # [nd*4H, F]
            b = torch.cat([(getattr(rnn, 'bias_ih_l0' + s) + getattr(rnn, 'bias_hh_l0' + s)).detach() for s in sfx], 0)
            whh = torch.stack([getattr(rnn, 'weight_hh_l0' + s).detach() for s in sfx], 0).contiguous()   # [nd,4H,H]
            geo = (B * S, K, 1, K, 0, 1) if which == 0 else (B * K, S, K, S * K, 1, K)
            hout, cst = (None if nohf else ops.empty(rows, nd * H)), ops.empty(rows, nd * H)
            # saved gate activations: fp32 row-major in the exact mode, bf16 packed per 8-unit chunk in the tensor-core mode
            gates = torch.empty((rows, nd * 4 * H), device=dev, dtype=torch.bfloat16 if ops.tf32 else torch.float32)
            if ops.tf32:
                # tensor-core recurrence (bf16 operands, fp32 accumulation and cell state), input projection fused:
                # the kernel of the inference path, whose epilogue also stores what BPTT needs
                from .engine import Engine
                xb = xb_next                                          # written by the previous half-block's norm + residual
                if xb is None:
                    xb = torch.empty((rows, F), device=dev, dtype=torch.bfloat16)
                    L_.call('dprnn_cast_bf16', xs, xb, rows * F, st)
                pp = model._engine.lstm_pingpong                      # half-job ping-pong kernel (same results)
                wp, bp = Engine._pack_lstm_tc(rnn, sfx, half_jobs=pp)
                hb = torch.empty((rows, nd * H), device=dev, dtype=torch.bfloat16)
                L_.call('dprnn_lstm_layer_bf16_train_pp' if pp else 'dprnn_lstm_layer_bf16_train', xb, wp, bp, hb, gates,
                        cst, hout, B, S, K, which, H, nd, int(model._engine.fast_act) | (_FWD_FLAGS if pp else 0), st)
                if not dg16:
                    xb = hb = None
            else:
                gx = ops.mm(xs, wih, rows, nd * 4 * H, F, bias=b)
                L_.call('dprnn_lstm_recurrence_f32_train', gx, whh.transpose(1, 2).contiguous(), hout, gates, cst, *geo, H, nd, st)
                del gx
            if nohf:
                yl = ops.mm16(hb, linm.weight, rows, F, nd * H, bias=linm.bias.detach())
            else:
                yl = ops.mm(hout, linm.weight.detach(), rows, F, nd * H, bias=linm.bias.detach())
            g_, b_, eps_ = _norm_params(nm)
            mr = ops.utt_stats(yl, B, S * K * F, eps_)
            n_half += 1
            xb_next = (torch.empty((rows, F), device=dev, dtype=torch.bfloat16)
                       if ops.tf32 and n_half < 2 * len(sep.dprnn_blocks) else None)      # the next LSTM's bf16 operand
            x_in = None
            if _KEEP_INPUTS and not dg16:       # x_out goes to a new buffer and the half-block's input stays for the backward (dW_ih)
                x_in, xs = xs, torch.empty_like(xs)
                L_.call('dprnn_norm_residual_to', yl, x_in, mr, g_.detach(), b_.detach(), B, S * K, F, xs, xb_next, st)
            else:                  # in place: the backward walks the (reversible) residual stream back, x_in = x_out - norm(y)
                L_.call('dprnn_norm_residual', yl, xs, mr, g_.detach(), b_.detach(), B, S * K, F, xb_next, st)
            halves.append(dict(nd=nd, geo=geo, hout=hout, gates=gates, cst=cst, yl=yl, mr=mr, wih=wih, whh=whh,
                               rnn=rnn, lin=linm, norm=nm, sfx=sfx, which=which, x_in=x_in,
                               xb=xb if dg16 else None, hb=hb if dg16 else None))
    c.update(halves=halves, xs=xs, dg16=dg16, nohf=nohf)

    # ---- PReLU, overlap-add, conv2d, gated head, end conv + activation (dprnn_spe.py:231-248)
    z = ops.empty(B, Lm, F)
    L_.call('dprnn_fold_prelu', xs, z, B, Lm, K, P, F, sep.prelu.weight.detach(), st)
    cw = sep.conv2d.weight.detach().reshape(2 * F, F)
    cb = sep.conv2d.bias.detach()
    wog = torch.cat([sep.out[0].weight.detach().reshape(F, F), sep.gate[0].weight.detach().reshape(F, F)], 0)   # [2F,F]
    bog = torch.cat([sep.out[0].bias.detach(), sep.gate[0].bias.detach()], 0)
    act = EPI_SIGMOID if cfg['activation_type'] == 'sigmoid' else EPI_RELU
    heads = []
    for sp in spks:
        u = ops.mm(z, cw[sp * F:(sp + 1) * F], B * Lm, F, F,
                   bias=(2.0 * cb[sp * F:(sp + 1) * F]).contiguous())         # overlap-add sums two chunks: bias twice
        pre = ops.mm(u, wog, B * Lm, 2 * F, F, bias=bog)
        g = ops.empty(B * Lm, F)
        L_.call('dprnn_gated_fwd', pre, g, B * Lm, F, st)
        m = ops.mm(g, sep.end_conv1x1.weight.detach().reshape(N, F), B * Lm, N, F, epi=act)
        heads.append(dict(sp=sp, u=u, pre=pre, g=g, m=m))
    c.update(z=z, heads=heads, wog=wog, cw=cw)
    return [hd['m'] for hd in heads], c


def forward_train(model, mix, ref=None, div=None, embedding=None):
    """-> est, logits, ctx (everything the backward needs).
    DPRNNTasNet (ref = div = None): est [B,2,T], logits None.  DPRNNSpeTasNet: est [B,T], logits [B,num_spks].
    DPRNNSpeIRATasNet: the two masker passes of dprnn_spe_ira.py:53-115.  embedding [B,E] (DPRNNRawNetTasNet, whose speaker
    encoder runs outside): replaces the speaker encoder; the backward then also returns its gradient."""
    _check_supported(model)
    L_, cfg, sep = lib(), model.cfg, model.separation
    dev = mix.device
    ops = _Ops(dev, tf32=model.precision == 'bf16')
    st = _st()
    N = cfg['input_size']
    B, T = mix.shape
    kind = cfg['kind']
    tss = kind != 'bss'
    Lm = T - 1
    w_enc = model.encoder.conv1d.weight.detach().reshape(N, 2).contiguous()
    w_dec = model.decoder.weight.detach().reshape(N, 2).contiguous()
    enc = ops.empty(B, Lm, N)
    L_.call('dprnn_encoder_fwd', mix, w_enc, enc, B, T, N, 2, 1, st)
    ctx = dict(B=B, T=T, L=Lm, mix=mix, ref=ref, enc=enc, tf32=ops.tf32, kind=kind, ext_emb=embedding is not None)
    emb = embedding
    if tss and embedding is None:
        Lr = ref.shape[1] - 1
        feats = ops.empty(B, Lr, N)
        L_.call('dprnn_encoder_fwd', ref, w_enc, feats, B, ref.shape[1], N, 2, 1, st)
        emb, ctx['spk0'] = _spk_fwd(model, ops, feats, B, Lr, div)
    _, _, eps = _norm_params(sep.bottleneck[0])
    mr_e = ops.utt_stats(enc, B, Lm * N, eps)
    ctx['mr_e'] = mr_e
    if kind == 'ira':
        # first pass, d0 = mask * enc, re-embedding (still divided by the REFERENCE's length, :84), aux_linear(cat(v0, v1))
        (m0,), ctx['core0'] = _core_fwd(model, ops, enc, mr_e, emb, B, Lm, (0,))
        d0 = ops.empty(B, Lm, N)
        L_.call('dprnn_mask_apply', m0, enc, d0, B * Lm * N, st)
        v1p, ctx['spk1'] = _spk_fwd(model, ops, d0, B, Lm, div)
        E = emb.shape[1]
        v = _emb_linear(ops, emb, sep.aux_linear, 0, E)                       # W[:, :E] v0 + b
        L_.call('dprnn_small_linear', v1p, E, sep.aux_linear.weight.detach().data_ptr() + 4 * E, 2 * E, None, v, E, B, E, E,
                1, st)                                                        # + W[:, E:] v1
        ctx.update(v0=emb, v1p=v1p)
        emb = v
    spks = (0,) if tss else (0, 1)
    masks, ctx['core'] = _core_fwd(model, ops, enc, mr_e, emb, B, Lm, spks)
    est = ops.empty(B, T) if tss else ops.empty(B, 2, T)
    for sp, m in zip(spks, masks):
        # DPRNN-Spe keeps speaker 0 only (dprnn_spe.py:325); DPRNN-TasNet decodes both (dprnn.py:277-281)
        L_.call('dprnn_mask_decode', m, Lm * N, enc, w_dec, est.data_ptr() + 4 * sp * T, len(spks) * T, B, Lm, N, 2, 1, st)
    logits = None
    if tss:
        logits = _emb_linear(ops, emb, sep.pred_linear)
    ctx['emb'] = emb
    return est, logits, ctx


# --------------------------------------------------------------------------------------------------------------
# backward
# --------------------------------------------------------------------------------------------------------------
def _emb_linear_bwd(ops, mod, name, emb, dv, demb, G, w_off=0, K=None):
    """dv [B, out]: gradient of mod(emb) (weight columns [w_off, w_off + K)) -> weight / bias gradients, demb += dv @ W"""
    B, E = emb.shape
    Kin = mod.weight.shape[1]
    K = E if K is None else K
    out = dv.shape[1]
    ops.atb(dv, emb, B, out, K, G[name + '.weight'].data_ptr() + 4 * w_off, ldb=E, ldc=Kin)
    if w_off == 0:
        ops.colsum(dv, B, out, G[name + '.bias'])
    wt = mod.weight.detach()[:, w_off:w_off + K].t().contiguous()       # [K, out]: demb[b,e] += sum_j dv[b,j] W[j,e]
    lib().call('dprnn_small_linear', dv, out, wt, out, None, demb, demb.shape[1], B, K, out, 1, _st())


def _spk_bwd(model, ops, sc, demb, G):
    """Adjoint of _spk_fwd: demb [B,E] -> gradient of its input features [B*Lr,N]; parameter gradients into G."""
    L_, sep, st, dev = lib(), model.separation, _st(), demb.device
    N = model.cfg['input_size']
    se = sep.spk_encoder
    B, Lr, L3, x3, div, feats = sc['B'], sc['Lr'], sc['L3'], sc['x3'], sc['div'], sc['feats']
    E = demb.shape[1]
    C5 = se[5].weight.shape[1]
    dscaled = torch.empty_like(demb)
    L_.call('dprnn_bcast_mul', (1.0 / div).contiguous().view(B, 1).expand(B, E).contiguous(), demb, dscaled, B, 1, E, 0, st)
    dz5 = ops.empty(B * L3, E)
    L_.call('dprnn_bcast_mul', dscaled, None, dz5, B, L3, E, 0, st)
    ops.atb(dz5, x3, B * L3, E, C5, G['separation.spk_encoder.5.weight'])
    ops.colsum(dz5, B * L3, E, G['separation.spk_encoder.5.bias'])
    dout = ops.mm(dz5, se[5].weight.detach().reshape(E, C5).t().contiguous(), B * L3, C5, E, x2=True)
    del dz5
    for bi, (rb, rc) in reversed(list(enumerate(zip((se[2], se[3], se[4]), sc['res'])))):
        pre_n = f'separation.spk_encoder.{bi + 2}'
        Cin, Cout, Lx, x = rc['Cin'], rc['Cout'], rc['Lx'], rc['x']
        rws = B * Lx
        # recompute v2 = BN2(y2) + skip and p2 = prelu(v2)
        v2 = ops.empty(rws, Cout)
        one = torch.ones(1, device=dev)
        L_.call('dprnn_affine_prelu', rc['y2'], rc['sc2'], rc['sh2'], one, v2, rws, Cout, st)      # slope 1 = identity
        ops.axpy(rc['skip'], v2)
        p2 = ops.empty(rws, Cout)
        ident_s, ident_b = torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)
        L_.call('dprnn_affine_prelu', v2, ident_s, ident_b, rb.prelu2.weight.detach(), p2, rws, Cout, st)
        dp2 = ops.empty(rws, Cout)
        L_.call('dprnn_pool3_bwd', dout, p2, dp2, B, Lx, Cout, st)
        dv2 = ops.prelu_bwd(dp2, v2, rb.prelu2.weight.detach(), G[pre_n + '.prelu2.weight'])
        del p2, dp2, v2

        def bn_bwd(dv, y, scale, shift, bnm, name):
            gmm = bnm.weight.detach()
            rstd = (scale / gmm).contiguous()
            mean = ((bnm.bias.detach() - shift) / scale).contiguous()
            s_d, s_dy = ops.empty(Cout), ops.empty(Cout)
            ops.colsum(dv, rws, Cout, s_d, accumulate=False)
            ops.colsum(dv, rws, Cout, s_dy, Y=y, accumulate=False)
            s_dyh = (rstd * (s_dy - mean * s_d)).contiguous()        # sum dv * yhat
            G[name + '.weight'] += s_dyh
            G[name + '.bias'] += s_d
            dy = ops.empty(rws, Cout)
            L_.call('dprnn_bn_bwd_apply', dv, y, mean, rstd, gmm, (s_d / rws).contiguous(), (s_dyh / rws).contiguous(), dy,
                    rws, Cout, st)
            return dy

        dy2 = bn_bwd(dv2, rc['y2'], rc['sc2'], rc['sh2'], rb.batch_norm2, pre_n + '.batch_norm2')
        ops.atb(dy2, rc['a1'], rws, Cout, Cout, G[pre_n + '.conv2.weight'])
        da1 = ops.mm(dy2, rb.conv2.weight.detach().reshape(Cout, Cout).t().contiguous(), rws, Cout, Cout, x2=True)
        v1 = ops.empty(rws, Cout)
        L_.call('dprnn_affine_prelu', rc['y1'], rc['sc1'], rc['sh1'], one, v1, rws, Cout, st)
        dv1 = ops.prelu_bwd(da1, v1, rb.prelu1.weight.detach(), G[pre_n + '.prelu1.weight'])
        dy1 = bn_bwd(dv1, rc['y1'], rc['sc1'], rc['sh1'], rb.batch_norm1, pre_n + '.batch_norm1')
        ops.atb(dy1, x, rws, Cout, Cin, G[pre_n + '.conv1.weight'])
        dxr = ops.mm(dy1, rb.conv1.weight.detach().reshape(Cout, Cin).t().contiguous(), rws, Cin, Cout, x2=True)
        if hasattr(rb, 'conv_downsample'):
            ops.atb(dv2, x, rws, Cout, Cin, G[pre_n + '.conv_downsample.weight'])
            t = ops.mm(dv2, rb.conv_downsample.weight.detach().reshape(Cout, Cin).t().contiguous(), rws, Cin, Cout, x2=True)
            ops.axpy(t, dxr)
        else:
            ops.axpy(dv2, dxr)
        dout = dxr
        del dy1, dy2, da1, dv1, dv2, v1
    O = se[1].weight.shape[0]
    ops.atb(dout, sc['gnf'], B * Lr, O, N, G['separation.spk_encoder.1.weight'])
    ops.colsum(dout, B * Lr, O, G['separation.spk_encoder.1.bias'])
    dgnf = ops.mm(dout, se[1].weight.detach().reshape(O, N).t().contiguous(), B * Lr, N, O, x2=True)
    return ops.gn_bwd(dgnf, feats, sc['mr_s'], se[0].weight.detach(), B, Lr, N, G['separation.spk_encoder.0.weight'],
                      G['separation.spk_encoder.0.bias'])


def _core_bwd(model, ops, c, dms, enc, mr_e, denc, demb, G):
    """Adjoint of _core_fwd.  dms: gradient of every mask it returned ([B*Lm,N] each); denc [B*Lm,N] receives (+=) the
    gradient reaching enc through the bottleneck norm; demb [B,E] (+=) the one reaching the embedding through the fusion."""
    L_, cfg, sep, st, dev = lib(), model.cfg, model.separation, _st(), enc.device
    N, F, H, K, P = cfg['input_size'], cfg['feature_size'], cfg['hidden_size'], cfg['chunk_length'], cfg['hop_length']
    B, Lm, S, rows, emb = c['B'], c['Lm'], c['S'], c['rows'], c['emb']
    E = emb.shape[1] if emb is not None else 0
    ML = B * Lm
    z = c['z']
    gout, ggate = G['separation.out.0.weight'], G['separation.gate.0.weight']
    gcw, gcb = G['separation.conv2d.weight'], G['separation.conv2d.bias']       # [2F,F,1,1]: rows of the decoded speakers
    dz = None
    for hd, dm in zip(c['heads'], dms):
        sp, m, g, pre, u = hd['sp'], hd['m'], hd['g'], hd['pre'], hd['u']
        dpm = torch.empty_like(dm)
        L_.call('dprnn_act_bwd', dm, m, dpm, dm.numel(), 2 if cfg['activation_type'] == 'sigmoid' else 1, st)
        ops.atb(dpm, g, ML, N, F, G['separation.end_conv1x1.weight'])
        dg = ops.mm(dpm, sep.end_conv1x1.weight.detach().reshape(N, F).t().contiguous(), ML, F, N)
        dpre = torch.empty_like(pre)
        L_.call('dprnn_gated_bwd', dg, pre, dpre, ML, F, st)
        ops.atb(dpre, u, ML, F, F, gout, lda=2 * F)
        ops.atb(dpre.data_ptr() + 4 * F, u, ML, F, F, ggate, lda=2 * F)
        ops.colsum(dpre, ML, F, G['separation.out.0.bias'], ldx=2 * F)
        ops.colsum(dpre.data_ptr() + 4 * F, ML, F, G['separation.gate.0.bias'], ldx=2 * F)
        du = ops.mm(dpre, c['wog'].t().contiguous(), ML, F, 2 * F)
        ops.atb(du, z, ML, F, F, gcw.data_ptr() + 4 * sp * F * F)
        dbc = ops.empty(F)
        ops.colsum(du, ML, F, dbc, accumulate=False)
        L_.call('dprnn_axpy', dbc, 2.0, gcb.data_ptr() + 4 * sp * F, F, 1, st)    # the folded conv adds the bias twice
        t = ops.mm(du, c['cw'][sp * F:(sp + 1) * F].t().contiguous(), ML, F, F)
        if dz is None:
            dz = t
        else:
            ops.axpy(t, dz)
        del du, dpre, dg, dpm, t
        hd.clear()
    # fold adjoint = unfold; then the PReLU adjoint on the final residual stream
    xs = c['xs']
    dxp = ops.empty(B, S, K, F)
    L_.call('dprnn_unfold', dz, dxp, B, Lm, K, P, F, st)
    dx = ops.prelu_bwd(dxp, xs, sep.prelu.weight.detach(), G['separation.prelu.weight'])
    del dxp, dz

    # ---- DPRNN blocks in reverse; the residual stream is walked back with x_in = x_out - norm(y)
    names = {}
    for n_, mod in model.named_modules():
        names[id(mod)] = n_
    # The gradient CHAIN (norm adjoint -> d h -> BPTT -> d x) runs on the caller's stream; the weight / bias gradients of each
    # half-block hang off it and run on a side stream, where they overlap the next half-block's BPTT (which occupies only
    # the SMs its tiles land on).  The residual stream is walked back OUT OF PLACE (x_in of a half-block goes to a new buffer),
    # so the chain never waits for the side stream's reads of the layer inputs; buffers handed to the side stream are
    # record_stream()-ed so that the caching allocator does not recycle them early.
    main = torch.cuda.current_stream()
    side = _side_stream(dev)
    side.wait_stream(main)

    def on_side(fn, *tensors):
        if _SKIP_SIDE:             # timing experiment only (tools/ab_cfg5.sh): the weight gradients are NOT computed
            return
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            fn()
        for t_ in tensors:
            t_.record_stream(side)

    for hv in reversed(c['halves']):
        nd, geo, yl, mr = hv['nd'], hv['geo'], hv['yl'], hv['mr']
        g_, b_, _ = _norm_params(hv['norm'])
        pn = names[id(hv['norm'])]
        gname = pn + ('.gamma' if hasattr(hv['norm'], 'gamma') else '.weight')
        bname = pn + ('.beta' if hasattr(hv['norm'], 'gamma') else '.bias')
        dg16 = c['dg16']
        if dg16:
            pass                                                     # the weight gradients read the kept bf16 input
        elif hv['x_in'] is not None:
            xs, hv['x_in'] = hv['x_in'], None                        # the half-block's input, kept by the forward
        else:
            xs_out, xs = xs, torch.empty_like(xs)
            L_.call('dprnn_norm_residual_to', yl, xs_out, mr, (-g_.detach()).contiguous(), (-b_.detach()).contiguous(), B,
                    S * K, F, xs, None, st)                          # xs: x_in = x_out - norm(y)
            del xs_out
        nohf = c['nohf']
        dy16 = torch.empty((rows, F), device=dev, dtype=torch.bfloat16) if nohf else None
        dy = ops.gn_bwd(dx, yl, mr, g_.detach(), B, S * K, F, G[gname], G[bname], dy16=dy16, only16=nohf)
        ln = names[id(hv['lin'])]
        hout = hv['hout']

        def lin_grads(dy=dy, hout=hout, ln=ln, nd=nd, dy16=dy16, hb=hv['hb']):
            if nohf:                       # bf16 operands: dW = dy^T [h_fwd | h_bwd], db = sum dy
                db = ops.empty(F)
                gw = G[ln + '.weight']
                if not ops.atb_dual(dy16, F, F, hb, nd * H, hb.data_ptr() + 2 * H, nd * H, B, S, K, 0, 0,
                                    gw, nd * H, gw.data_ptr() + 4 * H, nd * H, db, bf16=True):
                    raise RuntimeError('dprnn_gemm_atb_dual (bf16) does not apply to this shape')
                ops.axpy(db, G[ln + '.bias'])
                return
            if nd == 2 and F == 128:       # dW = dy^T [h_fwd | h_bwd] and db = sum dy from one pass over dy
                db = ops.empty(F)
                gw = G[ln + '.weight']
                if ops.atb_dual(dy, F, F, hout, nd * H, hout.data_ptr() + 4 * H, nd * H, B, S, K, 0, 0,
                                gw, nd * H, gw.data_ptr() + 4 * H, nd * H, db):
                    ops.axpy(db, G[ln + '.bias'])
                    return
            ops.atb(dy, hout, rows, F, nd * H, G[ln + '.weight'])
            ops.colsum(dy, rows, F, G[ln + '.bias'])
        on_side(lin_grads, *((dy16, hv['hb']) if nohf else (dy, hout)))
        if nohf:       # d h = dy W_lin with bf16 operands (dy exists as bf16 only)
            dh = ops.mm16(dy16, hv['lin'].weight.detach().t().contiguous(), rows, nd * H, F)
        else:
            dh = ops.mm(dy, hv['lin'].weight.detach().t().contiguous(), rows, nd * H, F)
        del dy
        dgates = torch.empty((rows, nd * 4 * H), device=dev, dtype=torch.bfloat16 if dg16 else torch.float32)
        if ops.tf32:
            whhT = hv['whh'].transpose(1, 2).contiguous().to(torch.bfloat16)              # [nd, H, 4H]
            L_.call('dprnn_lstm_bptt_tc_bf16out' if dg16 else 'dprnn_lstm_bptt_tc', dh, hv['gates'], hv['cst'], whhT, dgates,
                    *geo, H, nd, int(model._engine.fast_act) | _BPTT_FLAGS, st)
        else:
            L_.call('dprnn_lstm_bptt_f32', dh, hv['gates'], hv['cst'], hv['whh'], dgates, *geo, H, nd, st)
        del dh
        rn = names[id(hv['rnn'])]

        def rnn_grads(dgates=dgates, hout=hout, rn=rn, nd=nd, geo=geo, sfx=hv['sfx'], xs=xs, which=hv['which'],
                      xb=hv['xb'], hb=hv['hb']):
            if dg16:
                # bf16 operands: d gates as the BPTT's tensor-core tile held it, x and h as the forward's tensor cores read them
                for d, sf in enumerate(sfx):
                    db = ops.empty(4 * H)
                    if not ops.atb_dual(dgates.data_ptr() + 2 * d * 4 * H, nd * 4 * H, 4 * H, xb, F,
                                        hb.data_ptr() + 2 * d * H, nd * H, B, S, K, which, 1 if d else -1,
                                        G[f'{rn}.weight_ih_l0{sf}'], F, G[f'{rn}.weight_hh_l0{sf}'], H, db, bf16=True):
                        raise RuntimeError('dprnn_gemm_atb_dual (bf16) does not apply to this shape')
                    ops.axpy(db, G[f'{rn}.bias_ih_l0{sf}'])
                    ops.axpy(db, G[f'{rn}.bias_hh_l0{sf}'])
                return
            if F == 128 and ops.tf32 and ops.dual:
                # per direction ONE pass over its d gates: dW_ih = dg^T x, dW_hh = dg^T h_{t-1} (h read one time step earlier
                # - later for the reverse direction - through the tensor map: no shifted copy), db = sum dg
                done = True
                for d, sf in enumerate(sfx):
                    db = ops.empty(4 * H)
                    done = done and ops.atb_dual(dgates.data_ptr() + 4 * d * 4 * H, nd * 4 * H, 4 * H, xs, F,
                                                 hout.data_ptr() + 4 * d * H, nd * H, B, S, K, which, 1 if d else -1,
                                                 G[f'{rn}.weight_ih_l0{sf}'], F, G[f'{rn}.weight_hh_l0{sf}'], H, db)
                    if not done:
                        break
                    ops.axpy(db, G[f'{rn}.bias_ih_l0{sf}'])
                    ops.axpy(db, G[f'{rn}.bias_hh_l0{sf}'])
                if done:
                    return
            dbs = []
            for d, sf in enumerate(sfx):                             # the reads of xs first: the chain waits for them
                # dW_ih = dgates^T x; in tensor-core mode the same pass also yields db = column sums of dgates
                db = ops.empty(4 * H)
                if ops.atb_colsum(dgates.data_ptr() + 4 * d * 4 * H, xs, rows, 4 * H, F, G[f'{rn}.weight_ih_l0{sf}'], db,
                                  lda=nd * 4 * H):
                    dbs.append(db)
                else:
                    ops.atb(dgates.data_ptr() + 4 * d * 4 * H, xs, rows, 4 * H, F, G[f'{rn}.weight_ih_l0{sf}'], lda=nd * 4 * H)
                    dbs.append(None)
            hprev = ops.empty(rows, nd * H)
            L_.call('dprnn_shift_rows', hout, hprev, *geo, H, nd, _st())
            for d, sf in enumerate(sfx):
                dgd = dgates.data_ptr() + 4 * d * 4 * H
                ops.atb(dgd, hprev.data_ptr() + 4 * d * H, rows, 4 * H, H, G[f'{rn}.weight_hh_l0{sf}'], lda=nd * 4 * H, ldb=nd * H)
                db = dbs[d]                                          # b_ih and b_hh enter as a sum: one reduction, two adds
                if db is None:
                    db = ops.empty(4 * H)
                    ops.colsum(dgd, rows, 4 * H, db, ldx=nd * 4 * H, accumulate=False)
                ops.axpy(db, G[f'{rn}.bias_ih_l0{sf}'])
                ops.axpy(db, G[f'{rn}.bias_hh_l0{sf}'])
        on_side(rnn_grads, dgates, *((hv['xb'], hv['hb']) if dg16 else (xs, hout)))
        # dx (gradient of x_in) = dx_out + LSTM-branch gradient
        wihT = hv['wih'].t().contiguous()
        if dg16:
            if not ops.mm_acc(dgates, wihT.to(torch.bfloat16), rows, F, nd * 4 * H, dx):
                raise RuntimeError('dprnn_gemm_kdeep (bf16) does not apply to this shape')
        elif not ops.mm_acc(dgates, wihT, rows, F, nd * 4 * H, dx):
            dxl = ops.mm(dgates, wihT, rows, F, nd * 4 * H)
            ops.axpy(dxl, dx)
            del dxl
        del dgates, hout
        hv['hout'] = hv['gates'] = hv['cst'] = hv['yl'] = hv['xb'] = hv['hb'] = None       # free as we go
    main.wait_stream(side)

    # ---- unfold adjoint = fold; bottleneck conv; fusion; bottleneck norm
    dyb = ops.empty(B, Lm, F)
    L_.call('dprnn_fold_prelu', dx, dyb, B, Lm, K, P, F, None, st)
    del dx
    fused, mulc, addc = c['fused'], c['mulc'], c['addc']
    gbw = G['separation.bottleneck.1.weight']                       # [F, N(+E), 1]
    ldw = gbw.shape[1]
    ops.atb(dyb, fused, ML, F, N, gbw, ldc=ldw)
    ops.colsum(dyb, ML, F, G['separation.bottleneck.1.bias'])
    bw = sep.bottleneck[1].weight.detach().reshape(F, -1)
    dfused = ops.mm(dyb, bw[:, :N].t().contiguous(), ML, N, F)                    # first N input channels of the conv
    ft = cfg['fusion_type'] if emb is not None else None
    gamma, beta, _ = _norm_params(sep.bottleneck[0])
    bn0 = 'separation.bottleneck.0'
    gname = bn0 + ('.gamma' if hasattr(sep.bottleneck[0], 'gamma') else '.weight')
    bname = bn0 + ('.beta' if hasattr(sep.bottleneck[0], 'gamma') else '.bias')

    def lin_bwd(mod, name, dv):
        _emb_linear_bwd(ops, mod, name, emb, dv, demb, G)

    if ft == 'cat':
        # constant channels: y += W_e e per utterance -> de = sum_t dy @ W_e ; dW_e = (sum_t dy)^T e
        sdy = ops.utt_colsum(dyb, None, B, Lm, F)
        ops.atb(sdy, emb, B, F, E, gbw.data_ptr() + 4 * N, ldc=ldw)
        t = ops.gemm(sdy, bw.data_ptr() + 4 * N, B, E, F, ldw=ldw)
        ops.axpy(t, demb)
        dgn = dfused
    elif ft == 'att':
        # fused = n * v * r with r = 1 + softmax(<avg(n), v>)[src(l)]  (include/dprnn_b200.h, attention-fusion backward)
        ksz = cfg['kernel_size']
        s1, s0 = ops.empty(B, N), ops.empty(B, N)
        L_.call('dprnn_norm_affine', mr_e, gamma.detach(), beta.detach(), None, s1, s0, B, N, st)
        gn = torch.empty_like(enc)
        L_.call('dprnn_prologue_apply', enc, gn, ML, N, Lm, s1, s0, None, None, st)
        dr = ops.empty(B, Lm)
        L_.call('dprnn_row_dot3', dfused, gn, mulc, ML, Lm, N, dr, st)
        w2, ds = ops.empty(B, Lm), torch.empty_like(c['att_a'])
        L_.call('dprnn_att_softmax_bwd', dr, c['att_a'], B, Lm, ksz, w2, ds, st)
        dgn, tdv = torch.empty_like(dfused), torch.empty_like(dfused)
        L_.call('dprnn_att_bwd_apply', dfused, gn, mulc, c['rowscale'], w2, sep.average.weight.detach(), B, Lm, N, ksz,
                dgn, tdv, st)
        lin_bwd(sep.fusion_linear, 'separation.fusion_linear', ops.utt_colsum(tdv, None, B, Lm, N))
        del gn, tdv, dr, w2
    else:
        # fused = gn * mulc + addc (mulc / addc may be absent)
        if addc is not None:
            da2 = ops.utt_colsum(dfused, None, B, Lm, N)
            lin_bwd(sep.fusion_linear_2 if ft == 'film' else sep.fusion_linear, 'separation.fusion_linear_2' if ft == 'film'
                    else 'separation.fusion_linear', da2)
        if mulc is not None:
            # gn(enc) recomputed from the statistics
            s1, s0 = ops.empty(B, N), ops.empty(B, N)
            L_.call('dprnn_norm_affine', mr_e, gamma.detach(), beta.detach(), None, s1, s0, B, N, st)
            gn = torch.empty_like(enc)
            L_.call('dprnn_prologue_apply', enc, gn, ML, N, Lm, s1, s0, None, None, st)
            da1 = ops.utt_colsum(dfused, gn, B, Lm, N)
            lin_bwd(sep.fusion_linear_1 if ft == 'film' else sep.fusion_linear, 'separation.fusion_linear_1' if ft == 'film'
                    else 'separation.fusion_linear', da1)
            dgn = torch.empty_like(dfused)
            L_.call('dprnn_bcast_mul', mulc, dfused, dgn, B, Lm, N, 0, st)
            del gn
        else:
            dgn = dfused
    # GroupNorm(enc) adjoint, accumulated onto the gradient that reached enc through the mask product
    ops.gn_bwd(dgn, enc, mr_e, gamma.detach(), B, Lm, N, G[gname], G[bname], dy=denc, accumulate_dy=True)


def backward_train(model, ctx, d_est, d_logits, G=None):
    """-> {parameter name: gradient tensor} for every trainable parameter of the model.  ``G`` may be given (zeroed
    views into a flat gradient buffer, dp.FlatParams); every kernel accumulates into it.  With an external embedding
    (forward_train(embedding=...)) the gradient of that embedding is left in ctx['d_embedding']."""
    L_, cfg, sep = lib(), model.cfg, model.separation
    dev = d_est.device
    ops = _Ops(dev, tf32=ctx['tf32'])
    st = _st()
    N = cfg['input_size']
    B, T, Lm, enc, mr_e, kind = ctx['B'], ctx['T'], ctx['L'], ctx['enc'], ctx['mr_e'], ctx['kind']
    tss = kind != 'bss'
    if ctx.get('consumed'):
        raise RuntimeError('this forward has already been differentiated: the hand-written backward walks the saved '
                           'activations back in place (retain_graph / a second backward is not supported)')
    ctx['consumed'] = True
    if G is None:
        G = {n: torch.zeros_like(p) for n, p in model.named_parameters() if p.requires_grad}
    # frozen parameters (fine-tuning with e.g. a frozen speaker encoder; the reference optimises filter(requires_grad),
    # src/trainers/trainer.py:42-43): their gradients are computed into scratch and dropped
    G = dict(G)
    for n, p in model.named_parameters():
        if n not in G:
            G[n] = torch.zeros_like(p)
    d_est = d_est.contiguous().float()
    ML = B * Lm
    w_dec = model.decoder.weight.detach().reshape(N, 2).contiguous()
    wsw = torch.empty(L_.query('dprnn_convw2_workspace_bytes', N), device=dev, dtype=torch.uint8)
    encv = enc.view(ML, N)

    # ---- decoder and the mask product (per decoded speaker): dm = dz * enc, denc = sum dz * m
    core = ctx['core']
    denc, dms = None, []
    for hd in core['heads']:
        m = hd['m']
        de = d_est if tss else d_est[:, hd['sp']].contiguous()
        dze = ops.empty(B, Lm, N)
        L_.call('dprnn_decoder_bwd', de, w_dec, dze, B, Lm, N, st)
        me = ops.mul(m, encv)
        L_.call('dprnn_convw2_grad', me, de, B, Lm, N, G['decoder.weight'], 1, wsw, st)
        del me
        dms.append(ops.mul(dze.view(ML, N), encv))
        t = ops.mul(dze.view(ML, N), m)                             # gradient reaching enc through the mask product
        if denc is None:
            denc = t
        else:
            ops.axpy(t, denc)
        del dze
    emb = ctx['emb']
    demb = torch.zeros_like(emb) if tss else None
    _core_bwd(model, ops, core, dms, enc, mr_e, denc, demb, G)
    del dms
    dfeats = None
    if tss:
        _emb_linear_bwd(ops, sep.pred_linear, 'separation.pred_linear', emb, d_logits.contiguous().float(), demb, G)
    if kind == 'ira':
        # emb = aux_linear(cat(v0, v1')): gradients of both halves; v1' = speaker encoder of d0 = mask0 * enc
        v0, v1p, core0 = ctx['v0'], ctx['v1p'], ctx['core0']
        E = v0.shape[1]
        dv0, dv1p = torch.zeros_like(v0), torch.zeros_like(v1p)
        _emb_linear_bwd(ops, sep.aux_linear, 'separation.aux_linear', v0, demb, dv0, G, 0, E)
        _emb_linear_bwd(ops, sep.aux_linear, 'separation.aux_linear', v1p, demb, dv1p, G, E, E)
        dd0 = _spk_bwd(model, ops, ctx['spk1'], dv1p, G)             # [B*Lm, N]
        m0 = core0['heads'][0]['m']
        ops.axpy(ops.mul(dd0, m0), denc)
        dm0 = ops.mul(dd0, encv)
        del dd0
        _core_bwd(model, ops, core0, [dm0], enc, mr_e, denc, dv0, G)
        demb = dv0
    if tss and not ctx['ext_emb']:
        dfeats = _spk_bwd(model, ops, ctx['spk0'], demb, G)
    elif tss:
        ctx['d_embedding'] = demb

    # ---- encoder (shared by the mixture and the reference)
    genc = G['encoder.conv1d.weight']
    pairs = [(denc, enc, ctx['mix'])]
    if dfeats is not None:
        pairs.append((dfeats, ctx['spk0']['feats'], ctx['ref']))
    for d_, e_, sig in pairs:
        dpe = torch.empty_like(d_)
        L_.call('dprnn_act_bwd', d_, e_, dpe, d_.numel(), 1, st)
        L_.call('dprnn_convw2_grad', dpe, sig, B, e_.shape[1], N, genc, 1, wsw, st)
    return G


class TasNetTrainFunction(torch.autograd.Function):
    """The whole DPRNN-TasNet forward / backward as one autograd node over the model's trainable parameters."""

    @staticmethod
    def forward(fctx, model, mix, *params):
        est, _, ctx = forward_train(model, mix)
        fctx.model, fctx.saved = model, ctx
        fctx.names = [n for n, p in model.named_parameters() if p.requires_grad]
        return est

    @staticmethod
    def backward(fctx, d_est):
        G = backward_train(fctx.model, fctx.saved, d_est, None)
        fctx.saved = None
        return (None, None) + tuple(G[n] for n in fctx.names)


class SpeTrainFunction(torch.autograd.Function):
    """The whole DPRNN-Spe forward / backward as one autograd node over the model's trainable parameters."""

    @staticmethod
    def forward(fctx, model, mix, ref, div, *params):
        est, logits, ctx = forward_train(model, mix, ref, div)
        fctx.model, fctx.saved = model, ctx
        fctx.names = [n for n, p in model.named_parameters() if p.requires_grad]
        return est, logits

    @staticmethod
    def backward(fctx, d_est, d_logits):
        if d_logits is None:
            d_logits = torch.zeros((fctx.saved['B'], fctx.model.separation.pred_linear.weight.shape[0]), device=d_est.device)
        if d_est is None:
            d_est = torch.zeros((fctx.saved['B'], fctx.saved['T']), device=d_logits.device)
        G = backward_train(fctx.model, fctx.saved, d_est, d_logits)
        fctx.saved = None
        return (None, None, None, None) + tuple(G[n] for n in fctx.names)


class EmbTrainFunction(torch.autograd.Function):
    """Masker + decoder with an EXTERNAL speaker embedding (DPRNN-RawNet: RawNet3 runs outside, dprnn_rawnet.py:72-105) as
    one autograd node over the embedding and the model's trainable parameters other than the speaker encoder's."""

    @staticmethod
    def forward(fctx, model, mix, emb, *params):
        est, logits, ctx = forward_train(model, mix, embedding=emb.detach().contiguous().float())
        fctx.model, fctx.saved = model, ctx
        fctx.names = _core_param_names(model)
        return est, logits

    @staticmethod
    def backward(fctx, d_est, d_logits):
        if d_logits is None:
            d_logits = torch.zeros((fctx.saved['B'], fctx.model.separation.pred_linear.weight.shape[0]), device=d_est.device)
        if d_est is None:
            d_est = torch.zeros((fctx.saved['B'], fctx.saved['T']), device=d_logits.device)
        ctx = fctx.saved
        G = backward_train(fctx.model, ctx, d_est, d_logits)
        fctx.saved = None
        return (None, None, ctx['d_embedding']) + tuple(G[n] for n in fctx.names)


def _core_param_names(model):
    return [n for n, p in model.named_parameters() if p.requires_grad and not n.startswith('separation.spk_encoder.')]


def forward_with_grad(model, mix, ref=None, div=None, embedding=None):
    if embedding is not None:
        named = dict(model.named_parameters())
        return EmbTrainFunction.apply(model, mix, embedding, *[named[n] for n in _core_param_names(model)])
    params = [p for _, p in model.named_parameters() if p.requires_grad]
    if model.cfg['kind'] == 'bss':
        return TasNetTrainFunction.apply(model, mix, *params)
    return SpeTrainFunction.apply(model, mix, ref, div, *params)          # DPRNN-Spe and DPRNN-Spe-IRA


class SpeTrainStep:
    """One iteration of TrainerSpe.train (src/trainers/trainer_spe.py:27-56) - or, for DPRNNTasNet, of Trainer.train
    (src/trainers/trainer.py:100-118: PIT neg-SI-SDR, no speaker loss) - as device work only:
    zero_grad -> forward -> loss = mean neg-SI-SDR + ce_gamma * CE -> backward -> [all-reduce (mean) of the flat gradient
    buffer over the data-parallel group] -> clip_grad_norm_(max_norm) -> Adam(lr, weight_decay).

    The model's trainable parameters are re-seated as views into one flat fp32 buffer (dp.FlatParams), the backward
    accumulates straight into the matching flat gradient buffer, and the exchange step is ONE collective on it
    (SURVEY.md section 8e).  The reference has no data-parallel code; with world size 1 the step is the reference's."""

    def __init__(self, model, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, max_norm=5.0, ce_gamma=0.5,
                 group=None):
        from .dp import FlatParams, ClipAdam
        _check_supported(model)
        if model.cfg['kind'] == 'rawnet':
            raise NotImplementedError('DPRNN-RawNet trains through torch autograd (RawNet3 runs as library ops): use '
                                      'model(mix, ref16k) / loss.backward() with a torch optimiser, as TrainerRawNet does')
        if next(model.parameters()).device.type != 'cuda':
            raise RuntimeError('SpeTrainStep runs on the GPU (no CPU path)')
        model.train()
        self.model, self.group, self.ce_gamma = model, group, float(ce_gamma)
        self.fp = FlatParams(model)
        self.opt = ClipAdam(self.fp, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm)
        self.G = {n: p.grad for n, p in self.fp.named}
        dev = self.fp.flat.device
        self.loss3 = torch.zeros(3, device=dev)
        self.allreduce_events = None
        self.sync_replicas()

    def sync_replicas(self):
        """Data-parallel replicas must start from the same state (torch DDP does this at construction): broadcast rank
        0's parameters, Adam moments / step count and the speaker encoder's BatchNorm buffers.  The reference has no
        SyncBN, so the running statistics then evolve per rank on its own shard; average_bn_stats() (called by
        checkpoint()) puts the mean over the ranks into the file instead of rank 0's private copy."""
        from .dp import broadcast_state
        step = torch.tensor([float(self.opt.step_count)], device=self.fp.flat.device)
        broadcast_state([self.fp.flat, self.opt.exp_avg, self.opt.exp_avg_sq, step] + list(self.model.buffers()),
                        self.group)
        self.opt.step_count = int(step.item())
        self.model._engine.invalidate()

    def average_bn_stats(self):
        from .dp import average_buffers
        average_buffers([b for b in self.model.buffers() if b.dtype.is_floating_point], self.group)

    def loss_and_grads(self, mix, ref=None, target=None, spk_idx=None, ref_len=None):
        """forward + loss + backward into the flat gradient buffer (no exchange, no update).  -> loss3 (device tensor:
        total, SI-SDR part, CE part).  DPRNNSpeTasNet: (mix, ref, target [B,T], spk_idx [B]); DPRNNTasNet: (mix,
        target=targets [B,2,T]) with the two-source PIT assignment of Trainer (src/trainers/trainer.py:39,108-112)."""
        model = self.model
        mix, target = mix.contiguous().float(), target.contiguous().float()
        B, T = mix.shape
        self.fp.zero_grad()
        if model.cfg['kind'] == 'bss':
            est, _, ctx = forward_train(model, mix)
            tperm = torch.empty_like(target)
            self.perm = torch.empty(B, device=mix.device, dtype=torch.int32)
            lib().call('dprnn_pit2_assign', est, target, B, T, tperm, self.perm, None, _st())
            d_est = torch.empty_like(est)
            terms = torch.empty(2 * B, 2, device=mix.device)
            lib().call('dprnn_train_loss', est, tperm, T, None, 0, None, 0.0, 2 * B, terms, self.loss3, d_est, None, _st())
            backward_train(model, ctx, d_est, None, G=self.G)
            return self.loss3
        ref = ref.contiguous().float()
        div = model._engine._aux_div(ref.shape[1] if ref_len is None else ref_len, B, mix.device)
        est, logits, ctx = forward_train(model, mix, ref, div)
        C = logits.shape[1]
        d_est, d_logits = torch.empty_like(est), torch.empty_like(logits)
        terms = torch.empty(B, 2, device=mix.device)
        lib().call('dprnn_train_loss', est, target, T, logits, C, spk_idx.contiguous().long(), self.ce_gamma, B, terms,
                   self.loss3, d_est, d_logits, _st())
        backward_train(model, ctx, d_est, d_logits, G=self.G)
        return self.loss3

    def step(self, mix, ref=None, target=None, spk_idx=None, ref_len=None):
        from .dp import allreduce_mean
        loss3 = self.loss_and_grads(mix, ref, target, spk_idx, ref_len)
        if self.allreduce_events is not None:       # optional device timing of the exchange step (bench.py)
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        allreduce_mean(self.fp.grad, self.group)
        if self.allreduce_events is not None:
            ev[1].record()
            self.allreduce_events.append(ev)
        self.opt.step()
        self.model._engine.invalidate()       # the update went through raw pointers: cached weight packs are stale
        return loss3

    # ---- checkpoints in the reference's format (src/trainers/trainer.py:294-306: {'epoch', 'optimizer', 'model'})
    def checkpoint(self, epoch: int, average_bn: bool = False):
        """average_bn=True (a collective: call it on EVERY rank) stores the mean over the ranks of the BatchNorm
        running statistics instead of this rank's private ones."""
        if average_bn:
            self.average_bn_stats()
        return {'epoch': int(epoch), 'optimizer': self.opt.state_dict(),
                'model': {k: v.detach().clone() for k, v in self.model.state_dict().items()}}

    def save_checkpoint(self, path, epoch: int):
        torch.save(self.checkpoint(epoch), path)

    def load_checkpoint(self, cpt):
        """cpt: the dict (or a path to it) written by this class or by the reference's Trainer._save_checkpoint."""
        if not isinstance(cpt, dict):
            cpt = torch.load(cpt, map_location='cpu')
        with torch.no_grad():                                   # in place: the parameters stay views of the flat buffer
            own = self.model.state_dict()
            missing = set(own) - set(cpt['model'])
            if missing:
                raise KeyError(f'checkpoint lacks {sorted(missing)[:3]}...')
            for k, v in own.items():
                v.copy_(cpt['model'][k])
        self.model._engine.invalidate()
        self.opt.load_state_dict(cpt['optimizer'])
        self.sync_replicas()          # ranks that resumed from different files continue from rank 0's
        return cpt['epoch']


TrainStep = SpeTrainStep        # the same stepper drives both trainers (the model kind selects the loss)
