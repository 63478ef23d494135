"""Forward orchestration over the C ABI (include/dprnn_b200.h).

PyTorch is used here for device memory (``torch.empty``), the current CUDA stream and one-off weight
re-layout (transposes / concatenations of the ``state_dict`` tensors); every arithmetic stage of the
separation path is a kernel of libdprnn_b200.  No stage has a CPU or ATen fallback.

Internal layout is channels-last: frames ``[B, L, C]``, chunks ``[B, S, K, F]`` (a 128-float row per
chunk position), so intra-chunk sequences are contiguous rows and inter-chunk sequences a constant
row stride - see DESIGN.md.
"""
from __future__ import annotations

import os

import torch

from ._lib import lib
from .ragged import RaggedMixin

EPI_NONE, EPI_RELU, EPI_SIGMOID, EPI_GATED, EPI_AFFINE_PRELU = 0, 1, 2, 3, 4


def _t(w: torch.Tensor) -> torch.Tensor:
    """[N_out, K_in, ...1] weight -> contiguous [K_in, N_out]."""
    return w.detach().reshape(w.shape[0], -1).t().contiguous()


class Engine(RaggedMixin):
    def __init__(self, model):
        self.model = model
        self.precision = 'fp32'
        self.lstm_slices = 1           # time-sliced persistent LSTM kernel (uniform batches): 1 = off, k > 1 slices, 0 = auto
        self.lstm_pairs = 0            # > 0: cap on the CTA pairs that kernel keeps resident (0 = all SM pairs)
        # bf16 mode: the residual stream between the half-blocks lives in bf16 only (what the LSTM consumes anyway); the
        # fp32 master copy (False / env DPRNN_RESIDUAL_BF16=0) is ~1 dB closer to the reference and ~9 % slower
        self.residual_bf16 = os.environ.get('DPRNN_RESIDUAL_BF16', '1') == '1'
        self.lstm_pingpong = True      # half-job ping-pong LSTM kernel (bf16 mode, uniform batches; bit-identical results)
        self.fast_act = True       # tensor-core modes: tanh.approx-based gate activations (1 MUFU op each)
        # tensor-core modes, the 1x1 convolutions: None = by precision ('fp16' -> 'f32x2', 'bf16' -> 'tf32'); 'tf32' = fp32
        # operands truncated to TF32 by the tensor core; 'f32x2' = fp32 operands as bf16 pairs, 3 MMAs (16 significand bits:
        # TF32 truncation was the dominant error of the fp16 mode); 'fp32' = exact CUDA-core GEMM (diagnosis)
        self.conv_kind = None
        self.n_streams = 1         # >1: the batch is split into that many utterance groups on concurrent streams
        self.fused_tail = False    # bf16 mode: Linear + norm + residual as one persistent kernel (linear_norm.cu)
        # opt-in (tensor-core modes, 16-bit residual stream): the norm + residual of a half-block applied by the NEXT
        # half-block's LSTM kernel while it loads its input (lstm_tc_pp.cu, kFuse).  Bit-identical and one HBM pass less,
        # but measured SLOWER (intra 1.82 -> 4.04 ms per layer against a 0.40 ms norm pass saved: the two converter warps sit
        # on schedulers whose issue slots and MUFU pipe the cell updates already fill) - DESIGN.md section 4.2
        self.fuse_norm = os.environ.get('DPRNN_FUSE_NORM', '0') == '1'
        # tensor-core modes, 16-bit residual stream: Linear + norm + residual of a half-block as ONE launch whose Linear output
        # only lives in L2 (linear_normres.cu; bit-identical to the two kernels); 2 = also discard y from L2 after its use
        self.tail_l2 = int(os.environ.get('DPRNN_TAIL_L2', '0'))
        self.tail_lead = int(os.environ.get('DPRNN_TAIL_LEAD', '2'))    # utterances its Linear pass may run ahead of its norm pass
        # tensor-core modes, 16-bit residual stream: the last half-block's norm + residual is applied by the fold
        # (dprnn_norm_residual_fold_prelu_h16) and the unfold writes the 16-bit copy only, so the fp32 [B,S,K,F] tensor never
        # exists (3 x 1.59 GB of HBM traffic per forward at B = 64); bit-identical to the separate kernels (DPRNN_FOLD_FUSED=0)
        self.fold_fused = os.environ.get('DPRNN_FOLD_FUSED', '1') == '1'
        self._row_off = {}
        self._streams = []
        self.use_graphs = True     # eval forwards of a repeated shape are captured into a CUDA graph and replayed
        self._graphs = {}
        self._packed = None
        self._packed_key = None

    def set_precision(self, mode: str):
        """'fp32': exact fp32 on CUDA cores (parity mode).  'bf16' / 'fp16': the tcgen05 kernels with bf16 / fp16 operands
        (fp32 accumulation in TMEM, fp32 cell state and statistics).  'fp16' keeps the estimated sources within
        north_star's 1e-3 of the reference's fp32 path (11 significand bits instead of 8) at the speed of 'bf16'."""
        if mode not in ('fp32', 'bf16', 'fp16'):
            raise ValueError("precision must be 'fp32', 'bf16' or 'fp16'")
        if mode != 'fp32' and 'bf16' not in lib().build_info():
            raise RuntimeError('this build of libdprnn_b200 carries no tensor-core kernels')
        self.precision = mode

    @property
    def tc(self) -> bool:
        """tensor-core mode (bf16 or fp16 operands)"""
        return self.precision != 'fp32'

    @property
    def conv_mode(self) -> str:
        return getattr(self, 'conv_kind', None) or ('f32x2' if getattr(self, 'precision', 'bf16') == 'fp16' else 'tf32')

    @property
    def tc_conv(self) -> bool:
        return self.tc and self.conv_mode != 'fp32'

    @property
    def conv_tf32(self) -> bool:          # name kept for the call sites: "the convolutions run on the tensor cores"
        return self.conv_mode != 'fp32'

    def _w_x2(self, W):
        """[N, K] fp32 weight -> [N, 2K] bf16 for DPRNN_GEMM_F32X2: per 32 consecutive k, hi(32) then lo(32)."""
        cache = self.__dict__.setdefault('_x2', {})          # dropped by invalidate() together with the other packs
        key = (W.data_ptr(), tuple(W.shape))
        if key not in cache:
            N, K = W.shape
            w = W.detach().float().reshape(N, K // 32, 32)
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            cache[key] = (torch.cat([hi, lo], -1).reshape(N, 2 * K).contiguous(), W)     # W kept alive: its address is the key
        return cache[key][0]

    @property
    def h16(self) -> int:
        """DPRNN_H16_* code of the 16-bit operand / storage format"""
        return 1 if self.precision == 'fp16' else 0

    @property
    def h16_dtype(self):
        return torch.float16 if self.precision == 'fp16' else torch.bfloat16

    def _lstm_flags(self, s=None, which=0, ndir=2) -> int:
        """DPRNN_LSTM_FAST_ACT | DPRNN_LSTM_FP16 | tile size.  The launcher's own rule for 128-sequence tiles looks at ONE
        launch; with utterance groups on concurrent streams the layer of the whole batch is what shares the 74 CTA pairs,
        so the choice is made here from the batch of the forward (half tiles iff all its 128-sequence jobs fit one wave)."""
        f = int(self.fast_act) | (2 if self.precision == 'fp16' else 0)
        if s is not None:
            B = getattr(self, '_batch_total', None) or s['B']
            tiles = B * ((s['K'] + 127) // 128) if which else (B * s['S'] + 127) // 128
            f |= 4 if tiles * ndir <= 74 else 8          # DPRNN_LSTM_HALF_TILES / DPRNN_LSTM_FULL_TILES
        return f

    # ------------------------------------------------------------------ weights
    def invalidate(self):
        """Forget the kernel-layout weight copies and the captured graphs: for updates that write the parameters through
        raw pointers (the fused clip + Adam kernel), which do not bump the tensors' version counters."""
        self._packed = None
        self._packed_key = None
        self._graphs = {}
        self._x2 = {}

    def _weights_key(self):
        """Key of the kernel-layout weight copies: a re-seated or in-place-updated parameter changes it.  Writes torch
        cannot see (p.data, raw pointers) need model.invalidate()."""
        return (self.h16,) + tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def _graph_key(self):
        """Captured graphs additionally bake in the addresses (and, in eval mode, rely on the values) of the buffers -
        the BatchNorm running statistics."""
        return self._weights_key() + tuple((b.data_ptr(), b._version) for b in self.model.buffers())

    def packed(self):
        """Kernel-layout copies of the weights, rebuilt whenever a parameter changes."""
        key = self._weights_key()
        if self._packed is None or key != self._packed_key:
            self._x2 = {}
            self._packed = self._pack()
            self._packed_key = key
        return self._packed

    def _pack(self):
        m, cfg = self.model, self.model.cfg
        sep = m.separation
        N, F = cfg['input_size'], cfg['feature_size']
        W = {}
        W['enc'] = m.encoder.conv1d.weight.detach().reshape(N, -1).contiguous()
        W['dec'] = m.decoder.weight.detach().reshape(N, -1).contiguous()
        bw = sep.bottleneck[1].weight.detach().reshape(F, -1)           # [F, N(+E)]
        W['bott_wt'] = bw[:, :N].t().contiguous()
        W['bott_w_x'] = bw[:, :N].contiguous()          # [F, N] native layout for the tensor-core path
        W['bott_w_full'] = bw.contiguous()
        blocks = []
        for blk in sep.dprnn_blocks:
            halves = []
            for rnn, lin in ((blk.intra_rnn.rnn, blk.intra_linear), (blk.inter_rnn.rnn, blk.inter_linear)):
                sfx = ['', '_reverse'] if rnn.bidirectional else ['']
                wih = torch.cat([getattr(rnn, 'weight_ih_l0' + s).detach() for s in sfx], 0)       # [nd*4H, F]
                bias = torch.cat([(getattr(rnn, 'bias_ih_l0' + s) + getattr(rnn, 'bias_hh_l0' + s)).detach()
                                  for s in sfx], 0)
                whh = torch.stack([getattr(rnn, 'weight_hh_l0' + s).detach().t() for s in sfx], 0)  # [nd, H, 4H]
                wp, bp = self._pack_lstm_tc(rnn, sfx, dtype=self.h16_dtype)
                wp2, _ = self._pack_lstm_tc(rnn, sfx, half_jobs=True, dtype=self.h16_dtype)
                halves.append(dict(wih_t=wih.t().contiguous(), bias=bias.contiguous(), whh_t=whh.contiguous(),
                                   ndir=len(sfx), lin_t=_t(lin.weight), lin_b=lin.bias.detach(),
                                   tc_w=wp, tc_w2=wp2, tc_bias=bp, lin_bf16=lin.weight.detach().to(self.h16_dtype).contiguous()))
            blocks.append(halves)
        W['blocks'] = blocks
        cw = sep.conv2d.weight.detach().reshape(2 * F, F)
        W['conv2d_t'] = [cw[s * F:(s + 1) * F].t().contiguous() for s in range(2)]
        W['conv2d_b'] = [sep.conv2d.bias.detach()[s * F:(s + 1) * F].contiguous() for s in range(2)]
        # gated head: per 128-column tile, 64 'out' units followed by the matching 64 'gate' units
        wo, wg = sep.out[0].weight.detach().reshape(F, F), sep.gate[0].weight.detach().reshape(F, F)
        bo, bg = sep.out[0].bias.detach(), sep.gate[0].bias.detach()
        cols, bias = [], []
        for t0 in range(0, F, 64):
            cols += [wo[t0:t0 + 64].t(), wg[t0:t0 + 64].t()]
            bias += [bo[t0:t0 + 64], bg[t0:t0 + 64]]
        W['og_t'] = torch.cat(cols, 1).contiguous()       # [F, 2F]
        W['og_b'] = torch.cat(bias, 0).contiguous()
        W['end_t'] = _t(sep.end_conv1x1.weight)            # [F, N]
        # tensor-core (TF32) head: native [N_out, K] layouts
        W['conv2d_w'] = [cw[s * F:(s + 1) * F].contiguous() for s in range(2)]
        W['conv2d_b2'] = [(2.0 * sep.conv2d.bias.detach()[s * F:(s + 1) * F]).contiguous() for s in range(2)]
        W['og_w'] = torch.cat([wo, wg], 0).contiguous()     # [2F, F]: rows [out; gate]
        W['og_bias'] = torch.cat([bo, bg], 0).contiguous()
        W['end_w'] = sep.end_conv1x1.weight.detach().reshape(N, F).contiguous()
        if cfg['kind'] in ('spe', 'ira'):      # ResNet speaker encoder ('rawnet' carries RawNet3 instead)
            se = sep.spk_encoder
            W['spk_conv0_t'] = _t(se[1].weight)
            W['spk_res'] = [dict(c1=_t(rb.conv1.weight), c2=_t(rb.conv2.weight),
                                 down=_t(rb.conv_downsample.weight) if hasattr(rb, 'conv_downsample') else None)
                            for rb in (se[2], se[3], se[4])]
            W['spk_conv5_t'] = _t(se[5].weight)
        return W

    _lstm_perm_cache = {}

    @staticmethod
    def _pack_lstm_tc(rnn, sfx, half_jobs=False, dtype=torch.bfloat16):
        """Weight layout of dprnn_lstm_layer_bf16 (include/dprnn_b200.h): for direction d, CTA rank r and MMA
        instruction nh, the 128 rows {[W_ih | W_hh][q*H + 64*nh + j] : q in (2r, 2r+1), j < 64}; bias likewise.
        half_jobs: the layout of dprnn_lstm_layer_bf16_pp - rows {[..][gate*H + 64*nh + 32*r + u] : gate < 4, u < 32}."""
        H = rnn.hidden_size
        dev = rnn.weight_ih_l0.device
        key = (H, str(dev), bool(half_jobs))
        cache = Engine._lstm_perm_cache
        if key not in cache:        # row permutations and the 1/2 pre-scale are fixed: built once per device
            j = torch.arange(64)
            if half_jobs:
                u = torch.arange(32)
                wrows = torch.cat([torch.cat([g * H + 64 * nh + 32 * r + u for g in range(4)]) for r in range(2) for nh in range(2)])
            else:
                wrows = torch.cat([torch.cat([q * H + 64 * nh + j for q in (2 * r, 2 * r + 1)]) for r in range(2) for nh in range(2)])
            brows = torch.cat([q * H + 64 * nh + j for nh in range(2) for q in range(4)])
            half = torch.ones(4 * H)
            half[:2 * H] = 0.5      # i, f, o rows pre-scaled by 1/2 (exact): the kernel evaluates sigmoid(x) as
            half[3 * H:] = 0.5      # 1/2 tanh(x/2) + 1/2
            cache[key] = (wrows.to(dev), brows.to(dev), half.to(dev))
        wrows, brows, half = cache[key]
        ws, bs = [], []
        for sf in sfx:
            wcat = torch.cat([getattr(rnn, 'weight_ih_l0' + sf).detach(), getattr(rnn, 'weight_hh_l0' + sf).detach()], 1)
            b = (getattr(rnn, 'bias_ih_l0' + sf) + getattr(rnn, 'bias_hh_l0' + sf)).detach()
            ws.append((wcat * half[:, None])[wrows])
            bs.append((b * half)[brows])
        return torch.cat(ws, 0).to(dtype).contiguous(), torch.stack(bs, 0).float().contiguous()

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _check_input(self, x, name):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError(f'{name} must be a CUDA tensor: tss_with_dprnn_b200 has no CPU path')
        if x.dtype != torch.float32:
            raise TypeError(f'{name} must be float32')
        # every C-ABI call launches on the calling thread's CURRENT device and stream: a tensor or model on another
        # device would hand kernels of GPU a pointers of GPU b
        pdev = next(self.model.parameters()).device
        if x.device != pdev or x.device.index != torch.cuda.current_device():
            raise RuntimeError(f'{name} is on {x.device}, the model on {pdev}, the current CUDA device is '
                               f'cuda:{torch.cuda.current_device()}: run the call under torch.cuda.device(model device)')
        return x.contiguous()

    def _row_offsets(self, B, R, dev):
        """int64 [B+1] first chunk-position row of every utterance (equal-length batch)."""
        key = (B, R, str(dev))
        if key not in self._row_off:
            self._row_off[key] = (torch.arange(B + 1, dtype=torch.int64) * R).to(dev)
        return self._row_off[key]

    def _norm_params(self, mod):
        if hasattr(mod, 'gamma'):
            return mod.gamma.detach(), mod.beta.detach(), 1e-8       # GlobLN (norms.py:9)
        return mod.weight.detach(), mod.bias.detach(), mod.eps

    def _wants_grad(self):
        return torch.is_grad_enabled() and self.model.training and any(p.requires_grad for p in self.model.parameters())

    def _guard_autograd(self):
        if torch.is_grad_enabled() and self.model.training and any(p.requires_grad for p in self.model.parameters()):
            raise NotImplementedError('this entry point has no backward: call under torch.no_grad() or model.eval()')

    def gemm(self, A, Wt, M, N, K, out=None, bias=None, bias_per_utt=False, bias_scale=1.0, rows_per_utt=0,
             p_scale=None, p_shift=None, p_add=None, rowscale=None, epi=EPI_NONE):
        n_out = N // 2 if epi == EPI_GATED else N
        if out is None:
            out = torch.empty((M, n_out), device=A.device, dtype=torch.float32)
        lib().call('dprnn_gemm_f32', A, K, Wt, N, out, n_out, M, N, K, bias, int(bias_per_utt), float(bias_scale),
                   int(rows_per_utt), p_scale, p_shift, p_add, rowscale, epi, self._stream())
        return out

    def gemm_tc(self, A, W, M, N, K, bias=None, epi=EPI_NONE, out=None, stats=None, bias_rows_per_utt=0,
                bias_row_utt=None, post=None):
        """tcgen05 contraction; W in its native [N, K] layout (bf16 if A is bf16, else fp32 read as TF32).
        stats = (rows_per_utt, eps) additionally returns mean/rstd of the following per-utterance norm;
        bias_rows_per_utt / bias_row_utt select a per-utterance bias; post = (scale, shift, prelu_a) applies
        BatchNorm-eval affine + PReLU to the result (EPI_AFFINE_PRELU)."""
        L_ = lib()
        if post is not None:
            epi = EPI_AFFINE_PRELU
        n_out = N // 2 if epi == EPI_GATED else N
        if out is None:
            out = torch.empty((M, n_out), device=A.device, dtype=torch.float32)
        is_bf16 = int(A.dtype == torch.bfloat16)
        ps, psh, pa = post if post is not None else (None, None, None)
        if stats is None and A.dtype == torch.float32 and self.conv_mode == 'f32x2' and K % 32 == 0 and W.numel() == N * K:
            # fp32 operands as bf16 pairs (DPRNN_GEMM_F32X2); a weight too large to stay resident (N = K = 256: the last
            # ResBlock of the speaker encoder) is cut into two 128-row halves writing the two column halves of C
            parts = [(0, N)] if L_.query('dprnn_gemm_persist_supported', 2, N, K, epi) else \
                [(0, N // 2), (N // 2, N // 2)] if epi != EPI_GATED and L_.query('dprnn_gemm_persist_supported', 2, N // 2, K, epi) else []
            if parts and (len(parts) == 1 or not (bias_rows_per_utt or bias_row_utt is not None)):
                ws = torch.empty(L_.query('dprnn_gemm_persist_workspace_bytes'), device=A.device, dtype=torch.uint8)
                Wc = W.detach().reshape(N, K)          # a view: parameters and packed weights are contiguous
                for n0, nn in parts:
                    L_.call('dprnn_gemm_persist', A, 2, self._w_x2(Wc[n0:n0 + nn]), None if bias is None else bias[..., n0:] if bias.dim() == 1 else bias,
                            int(bias_rows_per_utt), bias_row_utt, None if ps is None else ps[n0:], None if psh is None else psh[n0:],
                            pa, out[:, n0:] if len(parts) > 1 else out, n_out, M, nn, K, epi, ws, self._stream())
                return out
        if stats is None and L_.query('dprnn_gemm_persist_supported', is_bf16, N, K, epi):
            ws = torch.empty(L_.query('dprnn_gemm_persist_workspace_bytes'), device=A.device, dtype=torch.uint8)
            L_.call('dprnn_gemm_persist', A, is_bf16, W, bias, int(bias_rows_per_utt), bias_row_utt, ps, psh, pa, out,
                    n_out, M, N, K, epi, ws, self._stream())
            return out
        if post is not None:
            L_.call('dprnn_gemm_tc_affine_prelu', A, W, post[0], post[1], post[2], out, n_out, M, N, K, self._stream())
            return out
        if bias_row_utt is not None:
            L_.call('dprnn_gemm_tc_ragged', A, is_bf16, W, bias, bias_row_utt, out, n_out, M, N, K, epi, self._stream())
            return out
        part = mr = None
        rpu, eps = int(bias_rows_per_utt), 0.0
        if stats is not None:
            rpu, eps = stats
            part = torch.empty(L_.query('dprnn_gemm_tc_stats_bytes', M), device=A.device, dtype=torch.uint8)
            mr = torch.empty((M // rpu, 2), device=A.device, dtype=torch.float32)
        L_.call('dprnn_gemm_tc', A, is_bf16, W, bias, out, n_out, M, N, K, epi, part,
                int(rpu), float(eps), mr, self._stream())
        return (out, mr) if stats is not None else out

    def utt_stats(self, x, B, elems, eps):
        L = lib()
        ws = torch.empty(L.query('dprnn_utt_stats_workspace_bytes', B), device=x.device, dtype=torch.uint8)
        mr = torch.empty((B, 2), device=x.device, dtype=torch.float32)
        L.call('dprnn_utt_stats', x, B, elems, float(eps), ws, mr, self._stream())
        return mr

    def small_linear(self, x, lin, B, out=None, accumulate=False, w_off=0, K=None, bias=True):
        Wm = lin.weight.detach()
        N, Kfull = Wm.shape[0], Wm.reshape(Wm.shape[0], -1).shape[1]
        K = Kfull - w_off if K is None else K
        if out is None:
            out = torch.empty((B, N), device=x.device, dtype=torch.float32)
        wptr = Wm.data_ptr() + 4 * w_off
        b = lin.bias.detach() if (bias and lin.bias is not None) else None
        lib().call('dprnn_small_linear', x, x.shape[1], wptr, Kfull, b, out, N, B, N, K, int(accumulate),
                   self._stream())
        return out

    def encode(self, wave):
        cfg, W = self.model.cfg, self.packed()
        B, T = wave.shape
        k, st, N = cfg['kernel_size'], cfg['stride'], cfg['input_size']
        if T < k:
            raise ValueError(f'input has {T} samples, fewer than the encoder kernel ({k})')
        L = (T - k) // st + 1
        enc = torch.empty((B, L, N), device=wave.device, dtype=torch.float32)
        lib().call('dprnn_encoder_fwd', wave, W['enc'], enc, B, T, N, k, st, self._stream())
        return enc, L

    # ------------------------------------------------------------------ speaker branch
    def _aux_div(self, aux_len, B, device):
        """aux_T of DPRNNSpe._auxiliary (dprnn_spe.py:159-160), from the caller's scalar (or [B]) tensor."""
        k = self.model.cfg['kernel_size']
        al = aux_len if isinstance(aux_len, torch.Tensor) else torch.tensor(float(aux_len))
        t = (al - k) // (k // 2) + 1
        t = ((t // 3) // 3) // 3
        t = t.reshape(-1).float().to(device)
        return t.expand(B).contiguous() if t.numel() == 1 else t.contiguous()

    def speaker_embedding(self, feats, B, Lr, div):
        """spk_encoder + time mean (dprnn_spe.py:115-122,156-163). feats [B,Lr,N] -> [B,E]."""
        L_, W, st = lib(), self.packed(), self._stream()
        se = self.model.separation.spk_encoder
        N = self.model.cfg['input_size']
        training = self.model.training
        dev = feats.device
        mr = self.utt_stats(feats, B, Lr * N, se[0].eps)
        s1 = torch.empty((B, N), device=dev); s0 = torch.empty_like(s1)
        L_.call('dprnn_norm_affine', mr, se[0].weight.detach(), se[0].bias.detach(), None, s1, s0, B, N, st)
        O = se[1].weight.shape[0]
        if self.tc_conv and N % 32 == 0 and O in (64, 128, 256):
            fn = torch.empty_like(feats)            # GroupNorm applied, then the 1x1 conv on the tensor cores (TF32)
            L_.call('dprnn_prologue_apply', feats, fn, B * Lr, N, Lr, s1, s0, None, None, st)
            x = self.gemm_tc(fn, se[1].weight.detach(), B * Lr, O, N, bias=se[1].bias.detach())
            del fn
        else:
            x = self.gemm(feats, W['spk_conv0_t'], B * Lr, O, N, bias=se[1].bias.detach(), rows_per_utt=Lr,
                          p_scale=s1, p_shift=s0)
        Lx = Lr
        for rb, wr in zip((se[2], se[3], se[4]), W['spk_res']):
            Cin, Cout = rb.conv1.weight.shape[1], rb.conv1.weight.shape[0]
            rows = B * Lx
            ws = torch.empty(L_.query('dprnn_bn_workspace_bytes', Cout), device=dev, dtype=torch.uint8)
            scale = torch.empty(Cout, device=dev); shift = torch.empty(Cout, device=dev)

            def bn(y, bnm):
                L_.call('dprnn_batchnorm_affine', y, rows, Cout, bnm.weight.detach(), bnm.bias.detach(),
                        bnm.running_mean, bnm.running_var, int(training), float(bnm.eps),
                        float(bnm.momentum if bnm.momentum is not None else 0.1), ws, scale, shift, st)
                if training:
                    bnm.num_batches_tracked += 1

            tc = self.tc_conv and Cin in (128, 256) and Cout in (128, 256)

            def conv(inp, conv_mod, wt, cin, cout):
                if tc:
                    return self.gemm_tc(inp, conv_mod.weight.detach(), rows, cout, cin)
                return self.gemm(inp, wt, rows, cout, cin)

            if tc and not training:
                # eval mode: the BatchNorm scale / shift are known before the conv -> conv + BN + PReLU in one pass
                bn(None, rb.batch_norm1)
                y = self.gemm_tc(x, rb.conv1.weight.detach(), rows, Cout, Cin,
                                 post=(scale, shift, rb.prelu1.weight.detach()))
            else:
                y = conv(x, rb.conv1, wr['c1'], Cin, Cout)
                bn(y, rb.batch_norm1)
                L_.call('dprnn_affine_prelu', y, scale, shift, rb.prelu1.weight.detach(), y, rows, Cout, st)
            y2 = conv(y, rb.conv2, wr['c2'], Cout, Cout)
            bn(y2, rb.batch_norm2)
            skip = x if wr['down'] is None else conv(x, rb.conv_downsample, wr['down'], Cin, Cout)
            Lo = Lx // 3
            out = torch.empty((B, Lo, Cout), device=dev)
            L_.call('dprnn_affine_add_prelu_pool3', y2, scale, shift, skip, rb.prelu2.weight.detach(), out, B, Lx,
                    Cout, st)
            x, Lx = out, Lo
        E = se[5].weight.shape[0]
        if self.tc_conv and E in (64, 128, 256) and se[5].weight.shape[1] % 32 == 0:
            z = self.gemm_tc(x, se[5].weight.detach(), B * Lx, E, se[5].weight.shape[1], bias=se[5].bias.detach())
        else:
            z = self.gemm(x, W['spk_conv5_t'], B * Lx, E, se[5].weight.shape[1], bias=se[5].bias.detach())
        emb = torch.empty((B, E), device=dev)
        L_.call('dprnn_time_sum', z, emb, B, Lx, E, div, st)
        return emb

    # ------------------------------------------------------------------ masker
    def masker(self, enc, mr, B, L, emb, speakers, outs=None):
        """bottleneck norm + fusion + 1x1 conv, segmentation, DPRNN blocks, PReLU, overlap-add, conv2d,
        gated head, activation (dprnn.py:166-187 / dprnn_spe.py:125-154,231-248).
        enc [B,L,N]; mr its GroupNorm statistics; returns one mask [B,L,N] per requested speaker."""
        s = self._masker_pre(enc, mr, B, L, emb)
        for bi in range(len(self.model.separation.dprnn_blocks)):
            for which in (0, 1):
                if s['bf16']:
                    self._half_lstm(s, bi, which)
                    self._half_tail(s, bi, which)
                else:
                    self._half_fp32(s, bi, which)
        return self._masker_post(s, speakers, outs)

    def _masker_pre(self, enc, mr, B, L, emb):
        """bottleneck norm + speaker fusion + 1x1 conv, segmentation, bf16 shadow; allocates the per-group buffers the
        blocks reuse (so that nothing is allocated or freed while two streams work on the group)."""
        L_, W, st = lib(), self.packed(), self._stream()
        cfg, sep = self.model.cfg, self.model.separation
        N, F, H = cfg['input_size'], cfg['feature_size'], cfg['hidden_size']
        K, P = cfg['chunk_length'], cfg['hop_length']
        dev = enc.device
        gamma, beta, _ = self._norm_params(sep.bottleneck[0])
        ft = cfg['fusion_type']
        mulc = addc = rowscale = None
        bias, bias_per_utt = sep.bottleneck[1].bias.detach(), False
        if ft == 'cat':       # constant channels -> a per-utterance bias W_e e + b
            bias = self.small_linear(emb, sep.bottleneck[1], B, w_off=N)
            bias_per_utt = True
        elif ft == 'add':
            addc = self.small_linear(emb, sep.fusion_linear, B)
        elif ft == 'mul':
            mulc = self.small_linear(emb, sep.fusion_linear, B)
        elif ft == 'film':
            mulc = self.small_linear(emb, sep.fusion_linear_1, B)
            addc = self.small_linear(emb, sep.fusion_linear_2, B)
        elif ft == 'att':
            mulc = self.small_linear(emb, sep.fusion_linear, B)
        s1 = torch.empty((B, N), device=dev); s0 = torch.empty_like(s1)
        if ft == 'att':
            k = cfg['kernel_size']
            n1 = torch.empty_like(s1); n0 = torch.empty_like(s1)
            L_.call('dprnn_norm_affine', mr, gamma, beta, None, n1, n0, B, N, st)
            La = (L - k) // k + 1
            scores = torch.empty((B, La), device=dev)
            rowscale = torch.empty((B, L), device=dev)
            L_.call('dprnn_att_rowscale', enc, n1, n0, sep.average.weight.detach(), sep.average.bias.detach(), mulc,
                    scores, rowscale, B, L, N, k, st)
        L_.call('dprnn_norm_affine', mr, gamma, beta, mulc, s1, s0, B, N, st)
        if self.tc_conv and N % 32 == 0 and F in (64, 128, 256):
            en = torch.empty_like(enc)              # norm + fusion applied, then the 1x1 conv on the tensor cores (TF32)
            L_.call('dprnn_prologue_apply', enc, en, B * L, N, L, s1, s0, addc, rowscale, st)
            y = self.gemm_tc(en, W['bott_w_x'], B * L, F, N, bias=bias, bias_rows_per_utt=L if bias_per_utt else 0)
            del en
        else:
            y = self.gemm(enc, W['bott_wt'], B * L, F, N, bias=bias, bias_per_utt=bias_per_utt, rows_per_utt=L,
                          p_scale=s1, p_shift=s0, p_add=addc, rowscale=rowscale)
        S = L_.query('dprnn_num_chunks', L, K, P)
        rows = B * S * K
        bf16 = self.tc
        # 16-bit residual stream (the default of the tensor-core modes): the fp32 [B,S,K,F] tensor never exists - the unfold
        # writes the 16-bit copy only and the last half-block's norm + residual is applied by the fold (_masker_post)
        x16_only = bool(bf16 and self.residual_bf16 and not self.fused_tail and self.fold_fused and len(sep.dprnn_blocks) > 0
                        and F % 8 == 0)
        x = None if x16_only else torch.empty((B, S, K, F), device=dev)
        s = dict(B=B, L=L, S=S, K=K, P=P, F=F, H=H, N=N, rows=rows, x=x, bf16=bf16, dev=dev)
        if bf16:
            if H != 128 or F != 128:
                raise NotImplementedError('the tensor-core LSTM kernel is built for feature_size = hidden_size = 128')
            s['xb'] = torch.empty((rows, F), device=dev, dtype=self.h16_dtype)
            L_.call('dprnn_unfold_h16', y, x, s['xb'], B, L, K, P, F, self.h16, st)
        else:
            L_.call('dprnn_unfold', y, x, B, L, K, P, F, st)
        del y
        if bf16:
            ndmax = max(hw['ndir'] for halves in W['blocks'] for hw in halves)
            s['hb'] = torch.empty((rows * ndmax * H,), device=dev, dtype=self.h16_dtype)
            s['ybuf'] = torch.empty((rows, F), device=dev, dtype=self.h16_dtype)
            s['part'] = torch.empty(L_.query('dprnn_gemm_tc_stats_bytes', rows), device=dev, dtype=torch.uint8)
            s['mr2'] = torch.empty((B, 2), device=dev)
            if self._fuses_norm():
                s['xb2'] = torch.empty_like(s['xb'])       # the residual stream ping-pongs between two 16-bit buffers
            if self.fused_tail:
                s['ws'] = torch.empty(L_.query('dprnn_linear_norm_workspace_bytes', rows, B), device=dev, dtype=torch.uint8)
        return s

    def _half_lstm(self, s, bi, which):
        """bf16 mode: one nn.LSTM layer (intra: which=0, inter: which=1) as the fused tcgen05 kernel -> s['hb']."""
        hw = self.packed()['blocks'][bi][which]
        pend = s.pop('pending_norm', None)
        if pend is not None:
            # the previous half-block left its norm + residual to this kernel: input = xb + norm(ybuf), applied while the
            # tiles are loaded; the updated residual stream lands in the other 16-bit buffer (lstm_tc_pp.cu, kFuse)
            g_, b_ = pend
            lib().call('dprnn_lstm_layer_bf16_pp_fused', s['xb'], s['ybuf'], s['mr2'], g_, b_, s['xb2'], hw['tc_w2'],
                       hw['tc_bias'], s['hb'], s['B'], s['S'], s['K'], which, s['H'], hw['ndir'], self._lstm_flags(),
                       self._stream())
            s['xb'], s['xb2'] = s['xb2'], s['xb']
            return
        if self.lstm_pingpong and self.lstm_slices != 1:
            # persistent CTA pairs over time-sliced jobs (lstm_tc_sliced.cu): the half-job ping-pong step, bit-identical
            # results, no part-empty last wave when the layer has more pair-jobs than the GPU has SM pairs
            key = ('lstm_ws', s['B'], s['S'], s['K'], which, hw['ndir'], torch.cuda.current_stream().cuda_stream)
            ws = s.get(key)
            if ws is None:
                ws = s[key] = torch.empty(lib().query('dprnn_lstm_sliced_workspace_bytes', s['B'], s['S'], s['K'], which,
                                                      hw['ndir']), device=s['dev'], dtype=torch.uint8)
            lib().call('dprnn_lstm_layer_bf16_sliced', s['xb'], hw['tc_w2'], hw['tc_bias'], s['hb'], s['B'], s['S'], s['K'],
                       which, s['H'], hw['ndir'], self._lstm_flags(), int(self.lstm_slices), int(self.lstm_pairs), ws,
                       self._stream())
            return
        if self.lstm_pingpong:
            # two half-jobs per CTA pair in ping-pong (lstm_tc_pp.cu): bit-identical results, the hand-off of one half-job
            # hidden under the cell update of the other
            lib().call('dprnn_lstm_layer_bf16_pp', s['xb'], hw['tc_w2'], hw['tc_bias'], s['hb'], s['B'], s['S'], s['K'],
                       which, s['H'], hw['ndir'], self._lstm_flags(s, which, hw['ndir']), self._stream())
            return
        if self.precision == 'fp16':
            raise NotImplementedError("precision 'fp16' is built for the default LSTM kernel (lstm_pingpong = True)")
        lib().call('dprnn_lstm_layer_bf16', s['xb'], hw['tc_w'], hw['tc_bias'], s['hb'], s['B'], s['S'], s['K'], which,
                   s['H'], hw['ndir'], int(self.fast_act), self._stream())

    def _half_tail(self, s, bi, which):
        """bf16 mode: Linear -> GroupNorm / gLN -> residual (dprnn.py:86-92 / 96-99) on s['hb'] -> s['x'], s['xb']."""
        if self.fused_tail:
            if self.precision == 'fp16':
                raise NotImplementedError("precision 'fp16' is built for the default two-kernel tail (fused_tail = False)")
            # Linear + norm statistics + norm apply + residual in one persistent kernel (linear_norm.cu)
            hw = self.packed()['blocks'][bi][which]
            blk = self.model.separation.dprnn_blocks[bi]
            g_, b_, eps = self._norm_params(blk.intra_norm if which == 0 else blk.inter_norm)
            R = s['S'] * s['K']
            lib().call('dprnn_linear_norm_residual_bf16', s['hb'], hw['lin_bf16'], hw['lin_b'], s['x'], s['xb'], g_, b_,
                       float(eps), self._row_offsets(s['B'], R, s['dev']), s['B'], R, s['rows'], hw['ndir'] * s['H'],
                       s['ws'], self._stream())
            return
        last = bi == len(self.model.separation.dprnn_blocks) - 1 and which == 1
        if self.tail_l2 and self.residual_bf16 and not last and not self._fuses_norm() and s['S'] * s['K'] >= 128:
            hw = self.packed()['blocks'][bi][which]
            blk = self.model.separation.dprnn_blocks[bi]
            g_, b_, eps = self._norm_params(blk.intra_norm if which == 0 else blk.inter_norm)
            if 'lnr_ws' not in s:
                s['lnr_ws'] = torch.empty(lib().query('dprnn_linear_normres_workspace_bytes', s['B']), device=s['dev'],
                                          dtype=torch.uint8)
            lib().call('dprnn_linear_normres_h16', s['hb'], hw['lin_bf16'], hw['lin_b'], s['ybuf'], s['xb'], g_, b_, s['rows'],
                       hw['ndir'] * s['H'], s['part'], s['S'] * s['K'], float(eps), s['mr2'], s['lnr_ws'],
                       int(self.tail_l2 == 2) | (int(self.tail_lead) << 8), self.h16, self._stream())
            return
        self._half_linear(s, bi, which)
        self._half_norm(s, bi, which)

    def _fuses_norm(self) -> bool:
        return (self.fuse_norm and self.tc and self.residual_bf16 and self.lstm_pingpong and self.lstm_slices == 1
                and not self.fused_tail)

    def _half_linear(self, s, bi, which):
        """Linear with bf16 output; per-utterance mean / rstd of the following norm from the fp32 accumulators."""
        hw = self.packed()['blocks'][bi][which]
        blk = self.model.separation.dprnn_blocks[bi]
        _, _, eps = self._norm_params(blk.intra_norm if which == 0 else blk.inter_norm)
        lib().call('dprnn_linear_h16out_stats', s['hb'], hw['lin_bf16'], hw['lin_b'], s['ybuf'], s['rows'],
                   hw['ndir'] * s['H'], s['part'], s['S'] * s['K'], float(eps), s['mr2'], self.h16, self._stream())

    def _half_norm(self, s, bi, which):
        blk = self.model.separation.dprnn_blocks[bi]
        g_, b_, _ = self._norm_params(blk.intra_norm if which == 0 else blk.inter_norm)
        last = bi == len(self.model.separation.dprnn_blocks) - 1 and which == 1     # nothing reads the bf16 shadow then
        if self._fuses_norm() and not last:
            s['pending_norm'] = (g_, b_)        # applied by the next half-block's LSTM kernel
            return
        if self.residual_bf16 and last and s['x'] is None:
            s['pending_fold'] = (g_, b_)        # applied by the fold: dprnn_norm_residual_fold_prelu_h16 (_masker_post)
            return
        if self.residual_bf16:      # residual stream in 16 bits only; the last half-block writes the fp32 x for the fold
            lib().call('dprnn_norm_residual_h16res', s['ybuf'], s['xb'], s['x'] if last else None, s['mr2'], g_, b_,
                       s['B'], s['S'] * s['K'], s['F'], self.h16, self._stream())
            return
        lib().call('dprnn_norm_residual_yh16', s['ybuf'], s['x'], s['mr2'], g_, b_, s['B'], s['S'] * s['K'], s['F'],
                   None if last else s['xb'], self.h16, self._stream())

    def _half_fp32(self, s, bi, which):
        """exact-fp32 mode: input projection GEMM, recurrence, Linear, statistics, norm + residual on CUDA cores."""
        L_, st = lib(), self._stream()
        hw = self.packed()['blocks'][bi][which]
        blk = self.model.separation.dprnn_blocks[bi]
        g_, b_, eps = self._norm_params(blk.intra_norm if which == 0 else blk.inter_norm)
        B, S, K, F, H, rows, x, nd = s['B'], s['S'], s['K'], s['F'], s['H'], s['rows'], s['x'], hw['ndir']
        gx = self.gemm(x, hw['wih_t'], rows, nd * 4 * H, F, bias=hw['bias'])
        hout = torch.empty((rows, nd * H), device=s['dev'])
        if which == 0:    # intra: one sequence per (b, s), steps along k
            geo = (B * S, K, 1, K, 0, 1)
        else:             # inter: one sequence per (b, k), steps along s
            geo = (B * K, S, K, S * K, 1, K)
        L_.call('dprnn_lstm_recurrence_f32', gx, hw['whh_t'], hout, geo[0], geo[1], geo[2], geo[3], geo[4],
                geo[5], H, nd, st)
        del gx
        yl = self.gemm(hout, hw['lin_t'], rows, F, nd * H, bias=hw['lin_b'])
        del hout
        mr2 = self.utt_stats(yl, B, S * K * F, eps)
        L_.call('dprnn_norm_residual', yl, x, mr2, g_, b_, B, S * K, F, None, st)

    def _masker_post(self, s, speakers, outs=None):
        """PReLU + overlap-add, conv2d (after the fold), gated head, end conv + activation -> masks."""
        L_, W, st = lib(), self.packed(), self._stream()
        cfg, sep = self.model.cfg, self.model.separation
        B, L, K, P, F, N, x, dev, bf16 = s['B'], s['L'], s['K'], s['P'], s['F'], s['N'], s['x'], s['dev'], s['bf16']
        z = torch.empty((B, L, F), device=dev)
        pend = s.pop('pending_fold', None)
        if pend is not None:        # last norm + residual + PReLU + overlap-add in one pass over the 16-bit tensors
            L_.call('dprnn_norm_residual_fold_prelu_h16', s['ybuf'], s['xb'], s['mr2'], pend[0], pend[1], z, B, L, K, P, F,
                    sep.prelu.weight.detach(), self.h16, st)
        else:
            L_.call('dprnn_fold_prelu', x, z, B, L, K, P, F, sep.prelu.weight.detach(), st)
        act = EPI_SIGMOID if cfg['activation_type'] == 'sigmoid' else EPI_RELU
        masks = []
        for spk in speakers:
            # conv2d after the fold: every frame is covered by exactly two chunks, hence 2*bias (A.6)
            cov = 2.0 if K == 2 * P else None
            if cov is None:
                raise NotImplementedError('hop_length must be chunk_length/2 (every shipped config)')
            out = None if outs is None else outs[len(masks)].view(B * L, N)
            if bf16 and self.conv_tf32 and F == 128 and N == 64:
                u = self.gemm_tc(z, W['conv2d_w'][spk], B * L, F, F, bias=W['conv2d_b2'][spk])
                g = self.gemm_tc(u, W['og_w'], B * L, 2 * F, F, bias=W['og_bias'], epi=EPI_GATED)
                masks.append(self.gemm_tc(g, W['end_w'], B * L, N, F, epi=act, out=out).view(B, L, N))
                continue
            u = self.gemm(z, W['conv2d_t'][spk], B * L, F, F, bias=W['conv2d_b'][spk], bias_scale=cov)
            g = self.gemm(u, W['og_t'], B * L, 2 * F, F, bias=W['og_b'], epi=EPI_GATED)
            masks.append(self.gemm(g, W['end_t'], B * L, N, F, out=out, epi=act).view(B, L, N))
        return masks

    def decode(self, mask, enc, out, B, L, out_utt_stride):
        cfg, W = self.model.cfg, self.packed()
        lib().call('dprnn_mask_decode', mask, L * cfg['input_size'], enc, W['dec'], out, out_utt_stride, B, L,
                   cfg['input_size'], cfg['kernel_size'], cfg['stride'], self._stream())

    # ------------------------------------------------------------------ whole-model forwards
    def _graphed(self, tag, inputs, fn):
        """fn(*inputs) -> tuple of tensors, replayed from a CUDA graph once the same (shape, weights, settings) has been
        seen before: a forward is ~400 kernel launches (x utterance groups), i.e. ~10 ms of host time that a graph
        replay does not pay - it matters most for the reference's B = 1 inference loop and for multi-stream batches.
        The first call of a key runs eagerly (one-off shapes never pay for a capture); the graph owns its buffers
        (static addresses: the TMA tensor maps baked into the captured launches stay valid); results are returned as
        fresh copies."""
        L_ = lib()
        if not self.use_graphs or self.model.training or L_.timing is not None or torch.cuda.is_current_stream_capturing():
            return fn(*inputs)
        key = (tag, tuple((tuple(t.shape), t.dtype, t.device.index) for t in inputs), self.precision, self.n_streams,
               self.fast_act, self.conv_mode, self.fused_tail, self.lstm_slices, self.lstm_pairs, self.lstm_pingpong, self.residual_bf16, self.fuse_norm, self.tail_l2, self.tail_lead, self.fold_fused, self._graph_key())
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs[key] = 'seen'
            while len(self._graphs) > 4:          # a graph pins every intermediate of its forward: keep few
                self._graphs.pop(next(iter(self._graphs)))
            return fn(*inputs)
        if ent == 'seen':
            static_in = [torch.empty_like(t) for t in inputs]
            for d, t in zip(static_in, inputs):
                d.copy_(t)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L_.launches
            with torch.cuda.graph(g):
                out = fn(*static_in)
            ent = (g, static_in, tuple(out), L_.launches - n0)
            L_.launches = n0
            self._graphs[key] = ent
        g, static_in, out, n_launch = ent
        for d, t in zip(static_in, inputs):
            d.copy_(t, non_blocking=True)
        g.replay()
        L_.launches += n_launch
        return tuple(o.clone() for o in out)

    def _run_groups(self, B, fn, couples_batch=False, alloc=None):
        """Run fn(b0, b1, outs) for utterance groups on concurrent streams.  alloc() -> the full-batch result tensors; every
        group writes its rows outs[i][b0:b1] in place (no concatenation pass afterwards).
        Utterances are independent (unless train-mode BatchNorm couples them), so the results do not depend on the
        grouping; side by side, one group's memory-bound and GEMM kernels (encoders, speaker ResNet, head) fill the
        SMs that another group's LSTM kernel leaves idle in its partial second wave."""
        n = 1 if couples_batch else max(1, min(self.n_streams, B))
        outs = alloc()
        self._batch_total = B
        if n == 1:
            fn(0, B, outs)
            return outs
        dev = torch.cuda.current_device()
        while len(self._streams) < n:
            self._streams.append(torch.cuda.Stream(device=dev))
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        base, extra = divmod(B, n)
        b0 = 0
        for i in range(n):
            b1 = b0 + base + (1 if i < extra else 0)
            st = self._streams[i]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                fn(b0, b1, tuple(o[b0:b1] for o in outs))
                done = torch.cuda.Event()
                done.record(st)
            main.wait_event(done)
            if not torch.cuda.is_current_stream_capturing():
                for t in outs:
                    t.record_stream(st)
            b0 = b1
        return outs

    def forward_bss(self, mix):
        mix = self._check_input(mix, 'input')
        cfg = self.model.cfg
        if cfg['kind'] == 'bss' and self._wants_grad():
            # training (scripts/train/config_bss.yaml): one autograd node over the hand-written forward / backward
            from .train import forward_with_grad
            return forward_with_grad(self.model, mix)
        self._guard_autograd()
        N = cfg['input_size']

        k, st_ = cfg['kernel_size'], cfg['stride']
        Tout = ((mix.shape[1] - k) // st_) * st_ + k

        def group_of(mixs, b0, b1, outs):
            m = mixs[b0:b1]
            B = b1 - b0
            enc, L = self.encode(m)
            _, _, eps = self._norm_params(self.model.separation.bottleneck[0])
            mr = self.utt_stats(enc, B, L * N, eps)
            masks = self.masker(enc, mr, B, L, None, (0, 1))
            for s in (0, 1):
                self.decode(masks[s], enc, outs[0][:, s], B, L, 2 * Tout)

        def alloc():
            return (torch.empty((mix.shape[0], 2, Tout), device=mix.device),)

        with torch.no_grad():
            return self._graphed('bss', (mix,), lambda m: self._run_groups(
                m.shape[0], lambda b0, b1, outs: group_of(m, b0, b1, outs), alloc=alloc))[0]

    def forward_spe(self, mix, ref, ref_len, embedding=None):
        mix = self._check_input(mix, 'input')
        if embedding is not None and self._wants_grad():
            # DPRNN-RawNet training: the speaker encoder ran outside (torch autograd); masker + decoder as one autograd node
            # that also returns the gradient of the embedding
            from .train import forward_with_grad
            return forward_with_grad(self.model, mix, embedding=self._check_input(embedding, 'embedding'))
        sep, cfg = self.model.separation, self.model.cfg
        N = cfg['input_size']

        k, st_ = cfg['kernel_size'], cfg['stride']
        Tout = ((mix.shape[1] - k) // st_) * st_ + k
        n_spk = sep.pred_linear.weight.shape[0]

        def group(m, r, d, e, b0, b1, outs):
            B = b1 - b0
            enc, L = self.encode(m[b0:b1])
            if e is None:
                feats, Lr = self.encode(r[b0:b1])
                emb = self.speaker_embedding(feats, B, Lr, d[b0:b1])
                del feats
            else:
                emb = e[b0:b1]
            _, _, eps = self._norm_params(sep.bottleneck[0])
            mr = self.utt_stats(enc, B, L * N, eps)
            mask = self.masker(enc, mr, B, L, emb, (0,))[0]
            self.decode(mask, enc, outs[0], B, L, Tout)
            self.small_linear(emb, sep.pred_linear, B, out=outs[1])

        def alloc():
            return (torch.empty((mix.shape[0], Tout), device=mix.device), torch.empty((mix.shape[0], n_spk), device=mix.device))

        if embedding is None and self._wants_grad():
            # training step (cfg 5): forward that keeps what the hand-written backward needs, as one autograd node
            from .train import forward_with_grad
            ref = self._check_input(ref, 'aux')
            return forward_with_grad(self.model, mix, ref, self._aux_div(ref_len, mix.shape[0], mix.device))
        with torch.no_grad():
            B = mix.shape[0]
            if embedding is None:
                ref = self._check_input(ref, 'aux')
                div = self._aux_div(ref_len, B, mix.device)
                # train-mode BatchNorm statistics run over the whole batch (dprnn_spe.py:20-21): no grouping then
                return self._graphed('spe', (mix, ref, div), lambda m, r, d: self._run_groups(
                    B, lambda b0, b1, outs: group(m, r, d, None, b0, b1, outs), couples_batch=self.model.training,
                    alloc=alloc))
            embedding = self._check_input(embedding, 'embedding')
            return self._graphed('spe_emb', (mix, embedding), lambda m, e: self._run_groups(
                B, lambda b0, b1, outs: group(m, None, None, e, b0, b1, outs), alloc=alloc))

    def forward_ira(self, mix, ref, ref_len):
        mix = self._check_input(mix, 'input')
        ref = self._check_input(ref, 'aux')
        if self._wants_grad():
            # training (dprnn_spe_ira.py:53-115 under TrainerSpe): both masker passes and both speaker-encoder passes as
            # one autograd node over the hand-written forward / backward
            from .train import forward_with_grad
            return forward_with_grad(self.model, mix, ref, self._aux_div(ref_len, mix.shape[0], mix.device))
        sep, cfg = self.model.separation, self.model.cfg
        N, E = cfg['input_size'], cfg['embeddings_size']

        k, st_ = cfg['kernel_size'], cfg['stride']
        Tout = ((mix.shape[1] - k) // st_) * st_ + k
        n_spk = sep.pred_linear.weight.shape[0]

        def group(m, r, d, b0, b1, outs):
            B = b1 - b0
            div = d[b0:b1]
            enc, L = self.encode(m[b0:b1])
            feats, Lr = self.encode(r[b0:b1])
            v0 = self.speaker_embedding(feats, B, Lr, div)
            del feats
            _, _, eps = self._norm_params(sep.bottleneck[0])
            mr = self.utt_stats(enc, B, L * N, eps)
            mask = self.masker(enc, mr, B, L, v0, (0,))[0]
            d0 = torch.empty_like(enc)
            lib().call('dprnn_mask_apply', mask, enc, d0, B * L * N, self._stream())
            v1 = self.speaker_embedding(d0, B, L, div)          # still divided by the reference's length (:84)
            del d0
            v = self.small_linear(v0, sep.aux_linear, B, K=E)                                  # W[:, :E] v0 + b
            self.small_linear(v1, sep.aux_linear, B, out=v, accumulate=True, w_off=E, bias=False)   # + W[:, E:] v1
            mask = self.masker(enc, mr, B, L, v, (0,))[0]
            self.decode(mask, enc, outs[0], B, L, Tout)
            self.small_linear(v, sep.pred_linear, B, out=outs[1])

        def alloc():
            return (torch.empty((mix.shape[0], Tout), device=mix.device), torch.empty((mix.shape[0], n_spk), device=mix.device))

        with torch.no_grad():
            B = mix.shape[0]
            div_all = self._aux_div(ref_len, B, mix.device)
            return self._graphed('ira', (mix, ref, div_all), lambda m, r, d: self._run_groups(
                B, lambda b0, b1, outs: group(m, r, d, b0, b1, outs), couples_batch=self.model.training, alloc=alloc))
