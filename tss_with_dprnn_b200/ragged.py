"""Ragged (variable-length) batches over the C ABI: full-length utterances of different durations packed back to
back (cfg 3).  The result for every utterance equals what the reference's B = 1 test loop
gives for it (src/inferencers/inferencer_spe.py:25-45, inferencer.py:54-71): no padding enters a norm statistic,
a recurrence, the attention softmax or the speaker time-mean.

Layouts (see csrc/ragged.cu): the *frame space* gives utterance b the rows [frame_off[b], frame_off[b] + T_b) - the
same indices as its samples in the packed waveform - of which the first L_b are frames; the *chunk space* gives it
S_b chunks of K rows.  Row-wise stages (1x1 convolutions, the intra-chunk LSTM, the Linear, casts) run on the packed
buffers with the uniform kernels; the per-utterance stages use the ``*_ragged`` entry points.
"""
from __future__ import annotations

import torch

from ._lib import lib

EPI_NONE, EPI_RELU, EPI_SIGMOID, EPI_GATED = 0, 1, 2, 3


class RaggedLayout:
    """Index arrays of one packed batch (host lists + device tensors)."""

    def __init__(self, lengths, cfg, device):
        k, st = cfg['kernel_size'], cfg['stride']
        if st != 1:
            raise NotImplementedError('ragged batches need encoder stride 1 (every shipped config: kernel 2, stride 1)')
        K, P = cfg['chunk_length'], cfg['hop_length']
        self.B = len(lengths)
        self.T = [int(t) for t in lengths]
        if min(self.T) < k:
            raise ValueError(f'an utterance is shorter than the encoder kernel ({k} samples)')
        self.L = [t - (k - 1) for t in self.T]
        self.S = [(l + K) // P + 1 for l in self.L]
        self.La = [(l - k) // k + 1 for l in self.L]
        self.K, self.k = K, k
        self.total_rows = sum(self.T)
        self.total_chunks = sum(self.S)
        fo, co = [0], [0]
        for t, s in zip(self.T, self.S):
            fo.append(fo[-1] + t)
            co.append(co[-1] + s)
        self.frame_off_h, self.chunk_off_h = fo, co
        i64 = dict(dtype=torch.int64, device=device)
        self.frame_off = torch.tensor(fo[:-1], **i64)
        self.chunk_off = torch.tensor(co[:-1], **i64)
        self.row_off = torch.tensor([c * K for c in co], **i64)            # [B+1] chunk-space rows
        self.L_d = torch.tensor(self.L, **i64)
        self.S_d = torch.tensor(self.S, **i64)
        self.La_d = torch.tensor(self.La, **i64)
        ar = torch.arange(self.B, dtype=torch.int32, device=device)
        self.frame_utt = torch.repeat_interleave(ar, torch.tensor(self.T, device=device), output_size=self.total_rows)
        self.chunk_utt = torch.repeat_interleave(ar, self.S_d, output_size=self.total_chunks)
        order = sorted(range(self.B), key=lambda b: -self.S[b])            # longest first (LPT)
        self.jobs = torch.tensor([[co[b], self.S[b]] for b in order], dtype=torch.int32, device=device)
        self.device = device
        self._pool = None

    def pool_stages(self):
        """Packed layouts of the three MaxPool1d(3) stages of the speaker ResNet (dprnn_spe.py:39-42)."""
        if self._pool is None:
            stages = []
            in_off, in_len = self.frame_off_h[:-1], self.L
            i64 = dict(dtype=torch.int64, device=self.device)
            for _ in range(3):
                out_len = [n // 3 for n in in_len]
                if min(out_len) < 1:
                    raise ValueError('an utterance is too short for the speaker encoder (needs >= 27 frames)')
                out_off = [0]
                for n in out_len:
                    out_off.append(out_off[-1] + n)
                out_len_d = torch.tensor(out_len, **i64)
                utt = torch.repeat_interleave(torch.arange(self.B, dtype=torch.int32, device=self.device), out_len_d,
                                              output_size=out_off[-1])
                stages.append(dict(in_off=torch.tensor(in_off, **i64), out_off=torch.tensor(out_off[:-1], **i64),
                                   out_len=out_len_d, out_utt=utt, total_out=out_off[-1]))
                in_off, in_len = out_off[:-1], out_len
            self._pool = stages
        return self._pool


class RaggedMixin:
    """Ragged counterparts of Engine.encode / speaker_embedding / masker / decode and the whole-model forwards."""

    # ------------------------------------------------------------------ packing
    def _pack_waves(self, waves, name):
        """list of 1-D waveforms, or an already packed (flat [sum T_b], lengths) pair -> (flat, RaggedLayout)."""
        if isinstance(waves, tuple) and len(waves) == 2 and isinstance(waves[0], torch.Tensor) and waves[0].dim() == 1 \
                and not isinstance(waves[1], torch.Tensor):
            flat, lengths = self._check_input(waves[0], name), [int(t) for t in waves[1]]
            if sum(lengths) != flat.numel():
                raise ValueError(f'{name}: the lengths do not add up to the packed waveform')
            return flat, self._layout(lengths, flat.device)
        if isinstance(waves, torch.Tensor) and waves.dim() == 2:
            waves = list(waves)
        ws = []
        for w in waves:
            w = self._check_input(w, name)
            if w.dim() == 2 and w.shape[0] == 1:
                w = w[0]
            if w.dim() != 1:
                raise ValueError(f'{name}: ragged batches take a list of 1-D waveforms')
            ws.append(w)
        return torch.cat(ws), self._layout([w.numel() for w in ws], ws[0].device)

    def _layout(self, lengths, device):
        """RaggedLayout of a batch, cached for the most recent length patterns (index arrays are pure functions of them)."""
        cache = self.__dict__.setdefault('_layouts', {})
        key = (tuple(lengths), str(device))
        lay = cache.pop(key, None)
        if lay is None:
            lay = RaggedLayout(lengths, self.model.cfg, device)
        cache[key] = lay                     # most recently used last
        while len(cache) > 64:
            cache.pop(next(iter(cache)))
        return lay

    def encode_ragged(self, flat, lay):
        cfg, W = self.model.cfg, self.packed()
        N, k = cfg['input_size'], cfg['kernel_size']
        enc = torch.empty((lay.total_rows, N), device=flat.device, dtype=torch.float32)
        if k > 1:
            enc[lay.total_rows - (k - 1):].zero_()       # rows past the last frame of the packed waveform
        lib().call('dprnn_encoder_fwd', flat, W['enc'], enc, 1, lay.total_rows, N, k, 1, self._stream())
        return enc

    def _stats_ragged(self, x, C, off, length, B, eps):
        L_ = lib()
        ws = torch.empty(L_.query('dprnn_utt_stats_ragged_workspace_bytes', B), device=x.device, dtype=torch.uint8)
        mr = torch.empty((B, 2), device=x.device, dtype=torch.float32)
        L_.call('dprnn_utt_stats_ragged', x, C, off, length, B, float(eps), ws, mr, self._stream())
        return mr

    def _gemm_ragged(self, A, Wt, M, N, K, row_utt, bias=None, bias_per_utt=False, p_scale=None, p_shift=None,
                     p_add=None, rowscale=None, epi=EPI_NONE):
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
        lib().call('dprnn_gemm_f32_ragged', A, K, Wt, N, out, N, M, N, K, bias, int(bias_per_utt), 1.0, row_utt,
                   p_scale, p_shift, p_add, rowscale, epi, self._stream())
        return out

    # ------------------------------------------------------------------ speaker branch
    def _aux_div_ragged(self, ref_lengths, device):
        """aux_T of DPRNNSpe._auxiliary (dprnn_spe.py:159-160) per utterance, from its own reference length."""
        k = self.model.cfg['kernel_size']
        al = torch.tensor([float(t) for t in ref_lengths])
        t = (al - k) // (k // 2) + 1
        t = ((t // 3) // 3) // 3
        return t.float().to(device).contiguous()

    def speaker_embedding_ragged(self, feats, lay, div):
        """spk_encoder + time mean (dprnn_spe.py:115-122,156-163) on a frame-space tensor [total_rows, N] -> [B,E].
        BatchNorm uses its running statistics: a packed batch reproduces per-utterance eval()-mode results."""
        if self.model.training:
            raise NotImplementedError('ragged batches reproduce eval()-mode results (train-mode BatchNorm statistics '
                                      'would couple the utterances of a batch); call model.eval() first')
        L_, W, st = lib(), self.packed(), self._stream()
        se = self.model.separation.spk_encoder
        N = self.model.cfg['input_size']
        B, dev = lay.B, feats.device
        mr = self._stats_ragged(feats, N, lay.frame_off, lay.L_d, B, se[0].eps)
        s1 = torch.empty((B, N), device=dev); s0 = torch.empty_like(s1)
        L_.call('dprnn_norm_affine', mr, se[0].weight.detach(), se[0].bias.detach(), None, s1, s0, B, N, st)
        O = se[1].weight.shape[0]
        if self.tc_conv and N % 32 == 0 and O in (64, 128, 256):     # same arithmetic as the uniform path
            fn = torch.empty_like(feats)
            L_.call('dprnn_prologue_apply_ragged', feats, fn, lay.total_rows, N, lay.frame_utt, s1, s0, None, None, st)
            x = self.gemm_tc(fn, se[1].weight.detach(), lay.total_rows, O, N, bias=se[1].bias.detach())
            del fn
        else:
            x = self._gemm_ragged(feats, W['spk_conv0_t'], lay.total_rows, O, N, lay.frame_utt, bias=se[1].bias.detach(),
                                  p_scale=s1, p_shift=s0)
        rows = lay.total_rows
        for rb, wr, stage in zip((se[2], se[3], se[4]), W['spk_res'], lay.pool_stages()):
            Cin, Cout = rb.conv1.weight.shape[1], rb.conv1.weight.shape[0]
            scale = torch.empty(Cout, device=dev); shift = torch.empty(Cout, device=dev)

            def bn(bnm):
                L_.call('dprnn_batchnorm_affine', None, rows, Cout, bnm.weight.detach(), bnm.bias.detach(),
                        bnm.running_mean, bnm.running_var, 0, float(bnm.eps), 0.1, None, scale, shift, st)

            tc = self.tc_conv and Cin in (128, 256) and Cout in (128, 256)

            def conv(inp, conv_mod, wt, cin, cout):
                if tc:
                    return self.gemm_tc(inp, conv_mod.weight.detach(), rows, cout, cin)
                return self.gemm(inp, wt, rows, cout, cin)

            if tc:          # same fused conv + BN + PReLU pass as the uniform eval path
                bn(rb.batch_norm1)
                y = self.gemm_tc(x, rb.conv1.weight.detach(), rows, Cout, Cin,
                                 post=(scale, shift, rb.prelu1.weight.detach()))
            else:
                y = conv(x, rb.conv1, wr['c1'], Cin, Cout)
                bn(rb.batch_norm1)
                L_.call('dprnn_affine_prelu', y, scale, shift, rb.prelu1.weight.detach(), y, rows, Cout, st)
            y2 = conv(y, rb.conv2, wr['c2'], Cout, Cout)
            bn(rb.batch_norm2)
            skip = x if wr['down'] is None else conv(x, rb.conv_downsample, wr['down'], Cin, Cout)
            out = torch.empty((stage['total_out'], Cout), device=dev)
            L_.call('dprnn_affine_add_prelu_pool3_ragged', y2, scale, shift, skip, rb.prelu2.weight.detach(), out,
                    stage['out_utt'], stage['in_off'], stage['out_off'], stage['total_out'], Cout, st)
            x, rows = out, stage['total_out']
        E = se[5].weight.shape[0]
        if self.tc_conv and E in (64, 128, 256) and se[5].weight.shape[1] % 32 == 0:
            z = self.gemm_tc(x, se[5].weight.detach(), rows, E, se[5].weight.shape[1], bias=se[5].bias.detach())
        else:
            z = self.gemm(x, W['spk_conv5_t'], rows, E, se[5].weight.shape[1], bias=se[5].bias.detach())
        last = lay.pool_stages()[-1]
        emb = torch.empty((B, E), device=dev)
        L_.call('dprnn_time_sum_ragged', z, emb, last['out_off'], last['out_len'], B, E, div, st)
        return emb

    # ------------------------------------------------------------------ masker
    def masker_ragged(self, enc, mr, lay, emb, speakers):
        """Engine.masker on a packed batch: enc [total_rows, N] (frame space) -> one mask [total_rows, N] per speaker."""
        L_, W, st = lib(), self.packed(), self._stream()
        cfg, sep = self.model.cfg, self.model.separation
        N, F, H = cfg['input_size'], cfg['feature_size'], cfg['hidden_size']
        K, P = cfg['chunk_length'], cfg['hop_length']
        if K != 2 * P:
            raise NotImplementedError('hop_length must be chunk_length/2 (every shipped config)')
        B, dev, TR = lay.B, enc.device, lay.total_rows
        gamma, beta, _ = self._norm_params(sep.bottleneck[0])
        ft = cfg['fusion_type']
        mulc = addc = rowscale = None
        bias, bias_per_utt = sep.bottleneck[1].bias.detach(), False
        if ft == 'cat':
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
            n1 = torch.empty_like(s1); n0 = torch.empty_like(s1)
            L_.call('dprnn_norm_affine', mr, gamma, beta, None, n1, n0, B, N, st)
            scores = torch.zeros(TR, device=dev)
            rowscale = torch.empty(TR, device=dev)
            L_.call('dprnn_att_rowscale_ragged', enc, n1, n0, sep.average.weight.detach(), sep.average.bias.detach(),
                    mulc, scores, rowscale, lay.frame_utt, lay.frame_off, lay.L_d, lay.La_d, B, TR, N,
                    cfg['kernel_size'], st)
        L_.call('dprnn_norm_affine', mr, gamma, beta, mulc, s1, s0, B, N, st)
        if self.tc_conv and N % 32 == 0 and F in (64, 128, 256):
            en = torch.empty_like(enc)
            L_.call('dprnn_prologue_apply_ragged', enc, en, TR, N, lay.frame_utt, s1, s0, addc, rowscale, st)
            y = self.gemm_tc(en, W['bott_w_x'], TR, F, N, bias=bias, bias_row_utt=lay.frame_utt if bias_per_utt else None)
            del en
        else:
            y = self._gemm_ragged(enc, W['bott_wt'], TR, F, N, lay.frame_utt, bias=bias, bias_per_utt=bias_per_utt,
                                  p_scale=s1, p_shift=s0, p_add=addc, rowscale=rowscale)
        TC = lay.total_chunks
        rows = TC * K
        bf16 = self.tc
        # 16-bit residual stream: the fp32 chunk-space tensor never exists - the unfold writes the 16-bit copy only and the last
        # half-block's norm + residual is applied by the fold (the uniform path's Engine.fold_fused, same arithmetic)
        x16_only = bool(bf16 and self.residual_bf16 and self.fold_fused and len(sep.dprnn_blocks) > 0)
        pending_fold = None
        x = None if x16_only else torch.empty((rows, F), device=dev)
        if bf16:
            if H != 128 or F != 128:
                raise NotImplementedError('the tensor-core LSTM kernel is built for feature_size = hidden_size = 128')
            if self.precision == 'fp16' and not self.lstm_pingpong:
                raise NotImplementedError("precision 'fp16' is built for the default LSTM kernel (lstm_pingpong = True)")
            xb = torch.empty((rows, F), device=dev, dtype=self.h16_dtype)
            L_.call('dprnn_unfold_ragged_h16', y, x, xb, lay.chunk_utt, lay.chunk_off, lay.frame_off, lay.L_d, TC, K, P, F,
                    self.h16, st)
        else:
            L_.call('dprnn_unfold_ragged', y, x, lay.chunk_utt, lay.chunk_off, lay.frame_off, lay.L_d, TC, K, P, F, st)
        del y
        for blk, halves in zip(sep.dprnn_blocks, W['blocks']):
            for which, hw in enumerate(halves):
                nd = hw['ndir']
                nm = blk.intra_norm if which == 0 else blk.inter_norm
                g_, b_, eps = self._norm_params(nm)
                mr2 = torch.empty((B, 2), device=dev)
                if bf16:
                    hb = torch.empty((rows, nd * H), device=dev, dtype=self.h16_dtype)
                    pp = '_pp' if self.lstm_pingpong else ''          # half-job ping-pong kernels (bit-identical)
                    wk = hw['tc_w2'] if self.lstm_pingpong else hw['tc_w']
                    if which == 0:      # every chunk is one length-K sequence: the packed chunk space is a uniform batch
                        L_.call('dprnn_lstm_layer_bf16' + pp, xb, wk, hw['tc_bias'], hb, 1, TC, K, 0, H, nd,
                                self._lstm_flags(), st)
                    else:               # one pair-job per utterance and direction, S_b steps each
                        L_.call('dprnn_lstm_inter_bf16_ragged' + pp, xb, wk, hw['tc_bias'], hb, TC, K, lay.jobs, B,
                                H, nd, self._lstm_flags(), st)
                    ybuf = torch.empty((rows, F), device=dev, dtype=self.h16_dtype)
                    part = torch.empty(L_.query('dprnn_gemm_tc_stats_bytes', rows), device=dev, dtype=torch.uint8)
                    self._linear_stats_ragged(hb, hw, ybuf, rows, nd * H, part, lay, eps, mr2)
                    if self.residual_bf16:      # opt-in bf16 residual stream (same arithmetic as the uniform path)
                        last = blk is sep.dprnn_blocks[len(sep.dprnn_blocks) - 1] and which == 1
                        if last and x is None:
                            pending_fold = (ybuf, mr2, g_, b_)      # applied by the fold below
                        else:
                            L_.call('dprnn_norm_residual_ragged_h16res', ybuf, xb, x if last else None, mr2, g_, b_,
                                    lay.chunk_utt, TC, K, F, self.h16, st)
                    else:
                        L_.call('dprnn_norm_residual_ragged', ybuf, 1 + self.h16, x, mr2, g_, b_, lay.chunk_utt, TC, K, F, xb, st)
                    del hb, ybuf
                    continue
                gx = self.gemm(x, hw['wih_t'], rows, nd * 4 * H, F, bias=hw['bias'])
                hout = torch.empty((rows, nd * H), device=dev)
                if which == 0:
                    L_.call('dprnn_lstm_recurrence_f32', gx, hw['whh_t'], hout, TC, K, 1, K, 0, 1, H, nd, st)
                else:                   # exact-fp32 mode: one launch per utterance (parity mode, not the fast path)
                    for b in range(B):
                        r0, Sb = lay.chunk_off_h[b] * K, lay.S[b]
                        L_.call('dprnn_lstm_recurrence_f32', gx[r0:], hw['whh_t'], hout[r0:], K, Sb, K, Sb * K, 1, K, H,
                                nd, st)
                del gx
                yl = self.gemm(hout, hw['lin_t'], rows, F, nd * H, bias=hw['lin_b'])
                del hout
                mr2 = self._stats_ragged(yl, F, lay.row_off[:-1], lay.S_d * K, B, eps)
                L_.call('dprnn_norm_residual_ragged', yl, 0, x, mr2, g_, b_, lay.chunk_utt, TC, K, F, None, st)
                del yl
        z = torch.empty((TR, F), device=dev)
        if pending_fold is not None:
            ybuf, mr2, g_, b_ = pending_fold
            L_.call('dprnn_norm_residual_fold_prelu_ragged_h16', ybuf, xb, mr2, g_, b_, z, lay.frame_utt, lay.frame_off,
                    lay.L_d, lay.chunk_off, lay.S_d, TR, K, P, F, sep.prelu.weight.detach(), self.h16, st)
            del pending_fold, ybuf
        else:
            L_.call('dprnn_fold_prelu_ragged', x, z, lay.frame_utt, lay.frame_off, lay.L_d, lay.chunk_off, lay.S_d, TR, K, P,
                    F, sep.prelu.weight.detach(), st)
        del x
        act = EPI_SIGMOID if cfg['activation_type'] == 'sigmoid' else EPI_RELU
        masks = []
        for spk in speakers:
            if bf16 and self.conv_tf32 and F == 128 and N == 64:
                u = self.gemm_tc(z, W['conv2d_w'][spk], TR, F, F, bias=W['conv2d_b2'][spk])
                g = self.gemm_tc(u, W['og_w'], TR, 2 * F, F, bias=W['og_bias'], epi=EPI_GATED)
                masks.append(self.gemm_tc(g, W['end_w'], TR, N, F, epi=act))
                continue
            u = self.gemm(z, W['conv2d_t'][spk], TR, F, F, bias=W['conv2d_b'][spk], bias_scale=2.0)
            g = self.gemm(u, W['og_t'], TR, 2 * F, F, bias=W['og_b'], epi=EPI_GATED)
            masks.append(self.gemm(g, W['end_t'], TR, N, F, epi=act))
        return masks

    def _linear_stats_ragged(self, hb, hw, ybuf, rows, Kdim, part, lay, eps, mr2):
        """Tensor-core Linear with bf16 output and per-row sums, then the per-utterance reduction."""
        L_, st = lib(), self._stream()
        L_.call('dprnn_linear_h16out_stats', hb, hw['lin_bf16'], hw['lin_b'], ybuf, rows, Kdim, part, 0, float(eps),
                None, self.h16, st)          # mean_rstd = NULL: per-row sums only
        L_.call('dprnn_row_stats_finalize_ragged', part, lay.row_off, lay.B, hw['lin_bf16'].shape[0], float(eps), mr2, st)

    def decode_ragged(self, mask, enc, lay):
        cfg, W = self.model.cfg, self.packed()
        out = torch.empty(lay.total_rows, device=enc.device)
        lib().call('dprnn_mask_decode_ragged', mask, enc, W['dec'], out, lay.frame_utt, lay.frame_off, lay.L_d,
                   lay.total_rows, cfg['input_size'], cfg['kernel_size'], self._stream())
        return out

    @staticmethod
    def _split(flat, lay):
        return [flat[lay.frame_off_h[b]:lay.frame_off_h[b + 1]] for b in range(lay.B)]

    # ------------------------------------------------------------------ whole-model forwards
    def _as_list(self, waves, name):
        """per-utterance 1-D views of a list / [B,T] tensor / packed (flat, lengths) pair"""
        if isinstance(waves, tuple) and len(waves) == 2 and isinstance(waves[0], torch.Tensor) and waves[0].dim() == 1 \
                and not isinstance(waves[1], torch.Tensor):
            flat = self._check_input(waves[0], name)
            lengths = [int(t) for t in waves[1]]
            if sum(lengths) != flat.numel():
                raise ValueError(f'{name}: the lengths do not add up to the packed waveform')
            return list(flat.split(lengths))
        return list(waves)

    def _ragged_groups(self, n_utt, run):
        """run(indices) -> (list of per-utterance tensors, [len(indices), C] tensor or None) for interleaved utterance
        groups on concurrent streams (as Engine._run_groups); results are re-assembled in the caller's order."""
        n = max(1, min(self.n_streams, n_utt))
        if n == 1:
            return run(list(range(n_utt)))
        while len(self._streams) < n:
            self._streams.append(torch.cuda.Stream(device=torch.cuda.current_device()))
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        groups = [list(range(g, n_utt, n)) for g in range(n)]
        results = []
        for g, idx in enumerate(groups):
            st = self._streams[g]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                res = run(idx)
                done = torch.cuda.Event()
                done.record(st)
            main.wait_event(done)
            for t in res[0]:
                t.record_stream(main)
            results.append(res)
        ests = [None] * n_utt
        second = None
        if results[0][1] is not None:
            second = torch.empty((n_utt,) + tuple(results[0][1].shape[1:]), device=results[0][1].device)
        for idx, (est, extra) in zip(groups, results):
            for j, b in enumerate(idx):
                ests[b] = est[j]
            if extra is not None:
                second[torch.tensor(idx, device=second.device)] = extra
        return ests, second

    def forward_bss_ragged(self, mixes):
        """list of [T_b] mixtures -> list of [2, T_b] estimates (DPRNNTasNet.forward per utterance, dprnn.py:271-283)."""
        mixes = self._as_list(mixes, 'input')
        with torch.no_grad():
            return self._ragged_groups(len(mixes), lambda idx: (self._bss_ragged_group([mixes[i] for i in idx]), None))[0]

    def forward_spe_ragged(self, mixes, refs, embedding=None):
        """lists of [T_b] mixtures and [Tr_b] references -> (list of [T_b] estimates, logits [B, num_spks]);
        DPRNNSpeTasNet.forward per utterance with aux_len = Tr_b (dprnn_spe.py:314-327, inferencer_spe.py:31-32)."""
        mixes = self._as_list(mixes, 'input')
        refs = None if embedding is not None else self._as_list(refs, 'aux')
        if refs is not None and len(refs) != len(mixes):
            raise ValueError('need one reference per mixture')
        with torch.no_grad():
            return self._ragged_groups(len(mixes), lambda idx: self._spe_ragged_group(
                [mixes[i] for i in idx], None if refs is None else [refs[i] for i in idx],
                None if embedding is None else embedding[torch.tensor(idx, device=embedding.device)]))

    def forward_ira_ragged(self, mixes, refs):
        """DPRNNSpeIRATasNet.forward per utterance (dprnn_spe_ira.py:53-115,179-190) on a packed batch."""
        mixes, refs = self._as_list(mixes, 'input'), self._as_list(refs, 'aux')
        if len(refs) != len(mixes):
            raise ValueError('need one reference per mixture')
        with torch.no_grad():
            return self._ragged_groups(len(mixes), lambda idx: self._ira_ragged_group([mixes[i] for i in idx],
                                                                                     [refs[i] for i in idx]))

    def _bss_ragged_group(self, mixes):
        with torch.no_grad():
            flat, lay = self._pack_waves(mixes, 'input')
            enc = self.encode_ragged(flat, lay)
            N = self.model.cfg['input_size']
            _, _, eps = self._norm_params(self.model.separation.bottleneck[0])
            mr = self._stats_ragged(enc, N, lay.frame_off, lay.L_d, lay.B, eps)
            masks = self.masker_ragged(enc, mr, lay, None, (0, 1))
            outs = [self._split(self.decode_ragged(m, enc, lay), lay) for m in masks]
            return [torch.stack([outs[0][b], outs[1][b]]) for b in range(lay.B)]

    def _spe_ragged_group(self, mixes, refs, embedding=None):
        with torch.no_grad():
            flat, lay = self._pack_waves(mixes, 'input')
            sep, cfg = self.model.separation, self.model.cfg
            N = cfg['input_size']
            enc = self.encode_ragged(flat, lay)
            if embedding is None:
                rflat, rlay = self._pack_waves(refs, 'aux')
                if rlay.B != lay.B:
                    raise ValueError('need one reference per mixture')
                feats = self.encode_ragged(rflat, rlay)
                emb = self.speaker_embedding_ragged(feats, rlay, self._aux_div_ragged(rlay.T, flat.device))
                del feats
            else:
                emb = self._check_input(embedding, 'embedding')
            _, _, eps = self._norm_params(sep.bottleneck[0])
            mr = self._stats_ragged(enc, N, lay.frame_off, lay.L_d, lay.B, eps)
            mask = self.masker_ragged(enc, mr, lay, emb, (0,))[0]
            est = self.decode_ragged(mask, enc, lay)
            logits = self.small_linear(emb, sep.pred_linear, lay.B)
            return self._split(est, lay), logits

    def _ira_ragged_group(self, mixes, refs):
        with torch.no_grad():
            flat, lay = self._pack_waves(mixes, 'input')
            rflat, rlay = self._pack_waves(refs, 'aux')
            if rlay.B != lay.B:
                raise ValueError('need one reference per mixture')
            sep, cfg = self.model.separation, self.model.cfg
            N, B = cfg['input_size'], lay.B
            enc = self.encode_ragged(flat, lay)
            feats = self.encode_ragged(rflat, rlay)
            div = self._aux_div_ragged(rlay.T, flat.device)
            v0 = self.speaker_embedding_ragged(feats, rlay, div)
            del feats
            _, _, eps = self._norm_params(sep.bottleneck[0])
            mr = self._stats_ragged(enc, N, lay.frame_off, lay.L_d, B, eps)
            mask = self.masker_ragged(enc, mr, lay, v0, (0,))[0]
            d0 = torch.empty_like(enc)
            lib().call('dprnn_mask_apply', mask, enc, d0, enc.numel(), self._stream())
            v1 = self.speaker_embedding_ragged(d0, lay, div)       # still divided by the reference's length (:84)
            del d0
            E = cfg['embeddings_size']
            v = self.small_linear(v0, sep.aux_linear, B, K=E)
            self.small_linear(v1, sep.aux_linear, B, out=v, accumulate=True, w_off=E, bias=False)
            mask = self.masker_ragged(enc, mr, lay, v, (0,))[0]
            est = self.decode_ragged(mask, enc, lay)
            logits = self.small_linear(v, sep.pred_linear, B)
            return self._split(est, lay), logits
