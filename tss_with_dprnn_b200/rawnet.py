"""RawNet3 speaker encoder of DPRNN-RawNet (src/models/rawnet/RawNet3.py, RawNetBasicBlock.py; cfg 4).

Parameter containers with the reference's module tree (so ``chkpts/dprnn-rawnet/*.pt`` strict-load and seeded
default initialisation reproduces the reference's weights), plus the forward.

The whole forward runs through the C ABI: the front-end (PreEmphasis, InstanceNorm,
parameterised sinc filterbank, log-abs, mean normalisation: csrc/rawnet.cu), the three Res2Net blocks (1x1 and
kernel-3 dilated convolutions as tensor-core contractions with fused bias / ReLU / BatchNorm / residual epilogues -
TF32 in bf16 mode, exact fp32 on CUDA cores in fp32 mode - max-pooling, AFMS), layer4, and the attentive statistics
pooling.  Channels-last [B*T, C] activations like the rest of the path; torch is only used for buffers and one-off weight
re-layout.  That is the eval() path (InferencerRawNet calls model.eval()); in train() mode - BatchNorm batch statistics, and
the backward TrainerRawNet needs - the branch runs as stock torch ops under autograd (embed_autograd) and only the masker it
feeds is hand-written.  There is no CPU path: the tensors must be CUDA tensors.

The sinc front-end restates ``asteroid_filterbanks==0.4.0`` ``ParamSincFB`` / ``Encoder`` (third-party, not vendored in
the reference, not installable here): parity for that part is UNPINNED.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from ._lib import lib


class _PreEmphasis(nn.Module):
    def __init__(self, coef: float = 0.97):
        super().__init__()
        self.register_buffer('flipped_filter', torch.FloatTensor([-coef, 1.0]).unsqueeze(0).unsqueeze(0))


class _ParamSincFB(nn.Module):
    """Parameters / buffers of asteroid_filterbanks.ParamSincFB(n_filters, 251, stride=10) at 16 kHz."""

    def __init__(self, n_filters, kernel_size, stride, sample_rate=16000.0, min_low_hz=50, min_band_hz=50):
        super().__init__()
        if kernel_size % 2 == 0:
            kernel_size += 1
        self.n_filters, self.kernel_size, self.stride = n_filters, kernel_size, stride
        self.sample_rate, self.min_low_hz, self.min_band_hz = sample_rate, min_low_hz, min_band_hz
        half = kernel_size // 2
        to_mel = lambda hz: 2595 * np.log10(1 + hz / 700)
        to_hz = lambda mel: 700 * (10 ** (mel / 2595) - 1)
        mel = np.linspace(to_mel(30.0), to_mel(sample_rate / 2 - (min_low_hz + min_band_hz)), n_filters // 2 + 1,
                          dtype='float32')
        hz = to_hz(mel)
        self.low_hz_ = nn.Parameter(torch.from_numpy(hz[:-1]).view(-1, 1))
        self.band_hz_ = nn.Parameter(torch.from_numpy(np.diff(hz)).view(-1, 1))
        self.register_buffer('window_', torch.from_numpy(np.hamming(kernel_size)[:half]).float())
        self.register_buffer('n_', 2 * np.pi * (torch.arange(-half, 0.0).view(1, -1) / sample_rate))

    def filters(self):
        low = self.min_low_hz + torch.abs(self.low_hz_)
        high = torch.clamp(low + self.min_band_hz + torch.abs(self.band_hz_), self.min_low_hz, self.sample_rate / 2)
        band = (high - low)[:, 0]
        ft_low, ft_high = torch.matmul(low, self.n_), torch.matmul(high, self.n_)
        cl = ((torch.sin(ft_high) - torch.sin(ft_low)) / (self.n_ / 2)) * self.window_
        sl = ((torch.cos(ft_low) - torch.cos(ft_high)) / (self.n_ / 2)) * self.window_
        cos_f = torch.cat([cl, 2 * band.view(-1, 1), torch.flip(cl, dims=[1])], dim=1)
        sin_f = torch.cat([sl, torch.zeros_like(band.view(-1, 1)), -torch.flip(sl, dims=[1])], dim=1)
        return (torch.cat([cos_f, sin_f], dim=0) / (2 * band.repeat(2)[:, None])).unsqueeze(1)


class _SincEncoder(nn.Module):
    def __init__(self, filterbank):
        super().__init__()
        self.filterbank = filterbank


class _AFMS(nn.Module):
    def __init__(self, nb_dim):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones((nb_dim, 1)))
        self.fc = nn.Linear(nb_dim, nb_dim)


class _Bottle2neck(nn.Module):
    """Container matching Bottle2neck (RawNetBasicBlock.py:58-109), same construction order."""

    def __init__(self, inplanes, planes, kernel_size, dilation, scale, pool):
        super().__init__()
        width = int(math.floor(planes / scale))
        self.conv1 = nn.Conv1d(inplanes, width * scale, kernel_size=1)
        self.bn1 = nn.BatchNorm1d(width * scale)
        self.nums = scale - 1
        convs, bns = [], []
        pad = math.floor(kernel_size / 2) * dilation
        for _ in range(self.nums):
            convs.append(nn.Conv1d(width, width, kernel_size=kernel_size, dilation=dilation, padding=pad))
            bns.append(nn.BatchNorm1d(width))
        self.convs = nn.ModuleList(convs)
        self.bns = nn.ModuleList(bns)
        self.conv3 = nn.Conv1d(width * scale, planes, kernel_size=1)
        self.bn3 = nn.BatchNorm1d(planes)
        self.width, self.dilation, self.pool = width, dilation, pool
        self.afms = _AFMS(planes)
        if inplanes != planes:
            self.residual = nn.Sequential(nn.Conv1d(inplanes, planes, kernel_size=1, stride=1, bias=False))
        else:
            self.residual = nn.Identity()


class RawNet3(nn.Module):
    """RawNet3(Bottle2neck, model_scale=8, context=True, summed=True, encoder_type='ECA', nOut=E, out_bn=False,
    sinc_stride=10, log_sinc=True, norm_sinc='mean') as DPRNNRawNet builds it (dprnn_rawnet.py:57-70)."""

    def __init__(self, nOut, C=1024, model_scale=8, sinc_stride=10):
        super().__init__()
        self.preprocess = nn.Sequential(_PreEmphasis(), nn.InstanceNorm1d(1, eps=1e-4, affine=True))
        self.conv1 = _SincEncoder(_ParamSincFB(C // 4, 251, stride=sinc_stride))
        self.bn1 = nn.BatchNorm1d(C // 4)           # registered by the reference, never used in its forward
        self.layer1 = _Bottle2neck(C // 4, C, 3, 2, model_scale, 5)
        self.layer2 = _Bottle2neck(C, C, 3, 3, model_scale, 3)
        self.layer3 = _Bottle2neck(C, C, 3, 4, model_scale, 0)
        self.layer4 = nn.Conv1d(3 * C, 1536, kernel_size=1)
        self.attention = nn.Sequential(nn.Conv1d(1536 * 3, 128, kernel_size=1), nn.ReLU(), nn.BatchNorm1d(128),
                                       nn.Conv1d(128, 1536, kernel_size=1), nn.Softmax(dim=2))
        self.bn5 = nn.BatchNorm1d(3072)
        self.fc6 = nn.Linear(3072, nOut)
        self.bn6 = nn.BatchNorm1d(nOut)             # out_bn=False: registered, unused
        self.allow_tf32 = False                     # contractions in full fp32 unless the model runs in bf16 mode (then TF32)

    # ------------------------------------------------------------------ hand-written kernel path
    def _packed(self):
        """Kernel-layout weights (rebuilt when a parameter / buffer changes): BatchNorm-eval scale / shift vectors,
        kernel-3 conv weights as [N, 3*C] (tap-major K), contiguous column slices of the attention conv."""
        key = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if getattr(self, '_pk', None) is not None and self._pk_key == key:
            return self._pk

        def bn(m):
            sc = (m.weight / torch.sqrt(m.running_var + m.eps)).detach().contiguous()
            return sc, (m.bias - m.running_mean * sc).detach().contiguous()

        def w2(conv):      # [N, C, k] -> [N, k*C]
            w = conv.weight.detach()
            return w.permute(0, 2, 1).reshape(w.shape[0], -1).contiguous()

        pk = {}
        for name in ('layer1', 'layer2', 'layer3'):
            blk = getattr(self, name)
            d = dict(w1=w2(blk.conv1), b1=blk.conv1.bias.detach(), bn1=bn(blk.bn1), w3=w2(blk.conv3),
                     b3=blk.conv3.bias.detach(), bn3=bn(blk.bn3),
                     convs=[(w2(c), c.bias.detach(), bn(b)) for c, b in zip(blk.convs, blk.bns)],
                     alpha=blk.afms.alpha.detach().reshape(-1).contiguous(),
                     wres=None if isinstance(blk.residual, nn.Identity) else w2(blk.residual[0]))
            pk[name] = d
        a = self.attention
        C4 = self.layer4.weight.shape[0]
        wa = a[0].weight.detach().reshape(a[0].weight.shape[0], -1)
        pk['w4'], pk['b4'] = w2(self.layer4), self.layer4.bias.detach()
        pk['wa_x'], pk['wa_m'], pk['wa_s'] = (wa[:, :C4].contiguous(), wa[:, C4:2 * C4].contiguous(),
                                              wa[:, 2 * C4:].contiguous())
        pk['bn_a'] = bn(a[2])
        pk['wa2'], pk['ba2'] = w2(a[3]), a[3].bias.detach()
        pk['bn5'] = bn(self.bn5)
        self._pk, self._pk_key = pk, key
        return pk

    def _gemm(self, A, W, M, N, K, out, ldc, col0=0, bias=None, relu_affine=None, residual=None, ldres=0, bias_rpu=0):
        """out[:, col0:col0+N] (row stride ldc) = A[M,K] @ W[N,K]^T (+ bias); relu_affine = (scale, shift) or (None, None)
        selects ReLU (+ BatchNorm affine) (+ residual) after the bias.  TF32 tensor cores in 256-column launches when the
        model runs in bf16 mode, exact fp32 on CUDA cores otherwise (W is then used through a cached transposed copy)."""
        L_, st = lib(), torch.cuda.current_stream().cuda_stream
        if not self.allow_tf32:
            cache = self.__dict__.setdefault('_wt', {})
            k = (W.data_ptr(), W._version)
            if k not in cache:
                cache[k] = W.t().contiguous()
            Wt = cache[k]
            cptr = out if isinstance(out, int) else out.data_ptr()
            cptr += 4 * col0
            if relu_affine is not None:
                sc, sh = relu_affine
                L_.call('dprnn_gemm_f32_relu_affine', A, K, Wt, N, bias, int(bias_rpu), sc, sh, residual, int(ldres), cptr,
                        int(ldc), M, N, K, st)
            else:
                L_.call('dprnn_gemm_f32', A, K, Wt, N, cptr, int(ldc), M, N, K, bias, 0, 1.0, 0, None, None, None, None, 0, st)
            return
        step = 256 if N % 256 == 0 else 128
        base = out if isinstance(out, int) else out.data_ptr()
        for n0 in range(0, N, step):
            Wn = W[n0:n0 + step]
            bn_ = None if bias is None else (bias[:, n0:n0 + step].contiguous() if bias_rpu else bias[n0:n0 + step])
            cptr = base + 4 * (col0 + n0)
            if relu_affine is not None:
                sc, sh = relu_affine
                res = None if residual is None else residual.data_ptr() + 4 * n0
                L_.call('dprnn_gemm_tc_relu_affine', A, Wn, bn_, int(bias_rpu), None if sc is None else sc[n0:n0 + step],
                        None if sh is None else sh[n0:n0 + step], res, int(ldres), cptr, int(ldc), M, step, K, st)
            else:
                L_.call('dprnn_gemm_tc', A, 0, Wn, bn_, cptr, int(ldc), M, step, K, 0, None, 0, 0.0, None, st)

    def _block(self, x, B, T, name, dil, pool, out=None, ldo=None):
        """Bottle2neck.forward (RawNetBasicBlock.py:111-142) on channels-last rows: x [B*T, Cin] -> [B*T/pool, planes]."""
        L_, st, pk = lib(), torch.cuda.current_stream().cuda_stream, self._packed()[name]
        blk = getattr(self, name)
        rows, dev = B * T, x.device
        Cin, planes, width = x.shape[1], pk['w3'].shape[0], blk.width
        if pk['wres'] is not None:
            res = torch.empty((rows, planes), device=dev)
            self._gemm(x, pk['wres'], rows, planes, Cin, res, planes)
        else:
            res = x
        o1 = torch.empty((rows, width * 8), device=dev)
        self._gemm(x, pk['w1'], rows, width * 8, Cin, o1, width * 8, bias=pk['b1'], relu_affine=pk['bn1'])
        col = torch.empty((rows, 3 * width), device=dev)
        for i, (w, b, bnp) in enumerate(pk['convs']):
            # sp = spx[i] (+ the previous conv's output), then conv(k=3, dilation) -> ReLU -> BN, written over spx[i]:
            # o1 becomes torch.cat((out_0 .. out_6, spx[7]), 1) in place
            a_ptr = o1.data_ptr() + 4 * width * i
            b_ptr = None if i == 0 else o1.data_ptr() + 4 * width * (i - 1)
            L_.call('dprnn_res2_gather', a_ptr, width * 8, b_ptr, width * 8, col, rows, T, width, dil, st)
            self._gemm(col, w, rows, width, 3 * width, o1, width * 8, col0=width * i, bias=b, relu_affine=bnp)
        o3 = torch.empty((rows, planes), device=dev)
        self._gemm(o1, pk['w3'], rows, planes, width * 8, o3, planes, bias=pk['b3'], relu_affine=pk['bn3'],
                   residual=res, ldres=planes)
        To = T
        if pool:
            To = T // pool
            p = torch.empty((B * To, planes), device=dev)
            L_.call('dprnn_maxpool_time', o3, None, p, planes, B, T, planes, pool, st)
            o3 = p
        mean = torch.empty((B, planes), device=dev)
        L_.call('dprnn_col_mean_std', o3, mean, None, B, To, planes, st)
        gate = torch.empty((B, planes), device=dev)
        fc = blk.afms.fc
        L_.call('dprnn_small_linear', mean, planes, fc.weight.detach(), planes, fc.bias.detach(), gate, planes, B, planes,
                planes, 0, st)
        L_.call('dprnn_affine_vec', gate, None, None, gate, B, planes, 1, st)
        if out is None:
            out, ldo = torch.empty((B * To, planes), device=dev), planes
        L_.call('dprnn_afms_apply', o3, pk['alpha'], gate, out, int(ldo), B, To, planes, st)
        return out, To

    def _embed_kernels(self, x):
        """RawNet3.forward (RawNet3.py:72-136) through the C ABI only.  x [B, T] raw 16 kHz reference -> [B, nOut]."""
        L_, st, pk = lib(), torch.cuda.current_stream().cuda_stream, self._packed()
        fb, inorm = self.conv1.filterbank, self.preprocess[1]
        x = x.contiguous()
        B, T = x.shape
        dev = x.device
        T0 = (T - fb.kernel_size) // fb.stride + 1
        feat = torch.empty((B * T0, fb.n_filters), device=dev)
        filt = torch.empty(fb.kernel_size * fb.n_filters, device=dev)
        stats = torch.empty(2 * B, device=dev)
        L_.call('dprnn_rawnet_frontend', x, B, T, inorm.weight.detach(), inorm.bias.detach(), fb.low_hz_.detach(),
                fb.band_hz_.detach(), fb.window_, fb.n_, fb.n_filters, fb.kernel_size, fb.stride, float(fb.sample_rate),
                filt, stats, feat, st)
        C = self.layer1.conv3.weight.shape[0]
        x1, T1 = self._block(feat, B, T0, 'layer1', 2, 5)
        T2 = T1 // 3
        cat = torch.empty((B * T2, 3 * C), device=dev)                       # torch.cat((mp3(x1), x2, x3), 1)
        L_.call('dprnn_maxpool_time', x1, None, cat, 3 * C, B, T1, C, 3, st)      # mp3(x1) -> cat[:, :C]
        x2, _ = self._block(x1, B, T1, 'layer2', 3, 3)
        L_.call('dprnn_maxpool_time', x2, None, cat.data_ptr() + 4 * C, 3 * C, B, T2, C, 1, st)      # copy -> cat[:, C:2C]
        x1p = torch.empty((B * T2, C), device=dev)
        L_.call('dprnn_maxpool_time', x1, None, x1p, C, B, T1, C, 3, st)
        s3 = torch.empty_like(x2)
        L_.call('dprnn_add2', x1p, x2, s3, s3.numel(), st)                   # summed=True (RawNet3.py:93)
        self._block(s3, B, T2, 'layer3', 4, 0, out=cat.data_ptr() + 4 * 2 * C, ldo=3 * C)
        C4 = self.layer4.weight.shape[0]
        rows = B * T2
        h = torch.empty((rows, C4), device=dev)
        self._gemm(cat, pk['w4'], rows, C4, 3 * C, h, C4, bias=pk['b4'], relu_affine=(None, None))
        mean = torch.empty((B, C4), device=dev); std = torch.empty((B, C4), device=dev)
        L_.call('dprnn_col_mean_std', h, mean, std, B, T2, C4, st)
        a = self.attention
        A0 = a[0].weight.shape[0]
        ab = torch.empty((B, A0), device=dev)                                # W_m mean + W_s std + b per utterance
        L_.call('dprnn_small_linear', mean, C4, pk['wa_m'], C4, a[0].bias.detach(), ab, A0, B, A0, C4, 0, st)
        L_.call('dprnn_small_linear', std, C4, pk['wa_s'], C4, None, ab, A0, B, A0, C4, 1, st)
        att = torch.empty((rows, A0), device=dev)
        self._gemm(h, pk['wa_x'], rows, A0, C4, att, A0, bias=ab, relu_affine=pk['bn_a'], bias_rpu=T2)
        logits = torch.empty((rows, C4), device=dev)
        self._gemm(att, pk['wa2'], rows, C4, A0, logits, C4, bias=pk['ba2'])
        pooled = torch.empty((B, 2 * C4), device=dev)
        L_.call('dprnn_att_stats_pool', h, logits, pooled, B, T2, C4, st)
        L_.call('dprnn_affine_vec', pooled, pk['bn5'][0], pk['bn5'][1], pooled, B, 2 * C4, 0, st)
        nOut = self.fc6.weight.shape[0]
        emb = torch.empty((B, nOut), device=dev)
        L_.call('dprnn_small_linear', pooled, 2 * C4, self.fc6.weight.detach(), 2 * C4, self.fc6.bias.detach(), emb, nOut,
                B, nOut, 2 * C4, 0, st)
        return emb

    # ------------------------------------------------------------------ training path (library ops + torch autograd)
    def _block_autograd(self, x, blk):
        """Bottle2neck.forward + AFMS (RawNetBasicBlock.py:48-55,111-142) with the module's own Conv1d / BatchNorm1d objects
        (BatchNorm honours self.training: batch statistics and running-stat updates in train mode). x [B,C,T]."""
        F = torch.nn.functional
        residual = blk.residual(x)
        out = blk.bn1(torch.relu(blk.conv1(x)))
        spx = torch.split(out, blk.width, 1)
        outs, sp = [], None
        for i in range(blk.nums):
            sp = spx[i] if i == 0 else sp + spx[i]
            sp = blk.bns[i](torch.relu(blk.convs[i](sp)))
            outs.append(sp)
        outs.append(spx[blk.nums])
        out = blk.bn3(torch.relu(blk.conv3(torch.cat(outs, 1)))) + residual
        if blk.pool:
            out = F.max_pool1d(out, blk.pool)
        y = torch.sigmoid(blk.afms.fc(out.mean(-1)))
        return (out + blk.afms.alpha) * y.unsqueeze(-1)

    def embed_autograd(self, x):
        """RawNet3.forward (RawNet3.py:72-136) as stock torch ops, differentiable and with train-mode BatchNorm - the
        TRAINING path of the speaker branch of DPRNN-RawNet (trainer_rawnet.py:31-56).  RawNet3 is ~8 % of the model's
        arithmetic; its hand-written kernels (embed) cover inference, and a hand-written
        backward for it is not built: here the branch runs on library kernels under torch autograd and hands its embedding
        to the hand-written masker forward / backward (train.EmbTrainFunction), which returns d loss / d embedding to it."""
        F = torch.nn.functional
        if not x.is_cuda:
            raise RuntimeError('aux must be a CUDA tensor: tss_with_dprnn_b200 has no CPU path')
        fb = self.conv1.filterbank
        xi = F.conv1d(F.pad(x.unsqueeze(1), (1, 0), 'reflect'), self.preprocess[0].flipped_filter)      # PreEmphasis
        xi = self.preprocess[1](xi)                                                                       # InstanceNorm1d
        f = torch.log(torch.abs(F.conv1d(xi, fb.filters(), stride=fb.stride)) + 1e-6)                     # RawNet3.py:79-81
        f = f - f.mean(-1, keepdim=True)                                                                  # :83
        x1 = self._block_autograd(f, self.layer1)
        x2 = self._block_autograd(x1, self.layer2)
        x1p = F.max_pool1d(x1, 3)
        x3 = self._block_autograd(x1p + x2, self.layer3)                                                  # summed (:93)
        h = torch.relu(self.layer4(torch.cat((x1p, x2, x3), 1)))
        t = h.shape[-1]
        g = torch.cat((h, h.mean(2, keepdim=True).repeat(1, 1, t),
                       torch.sqrt(h.var(2, keepdim=True).clamp(min=1e-4, max=1e4)).repeat(1, 1, t)), 1)  # :105-117
        w = self.attention(g)
        mu = torch.sum(h * w, 2)
        sg = torch.sqrt((torch.sum(h ** 2 * w, 2) - mu ** 2).clamp(min=1e-4, max=1e4))
        return self.fc6(self.bn5(torch.cat((mu, sg), 1)))                                                 # out_bn=False

    @torch.no_grad()
    def embed(self, x):
        """RawNet3.forward (RawNet3.py:72-136), eval mode, on the GPU.  x [B, T] raw 16 kHz reference -> [B, nOut]."""
        if not x.is_cuda:
            raise RuntimeError('aux must be a CUDA tensor: tss_with_dprnn_b200 has no CPU path')
        if self.training:
            return self.embed_autograd(x)          # train-mode BatchNorm (batch statistics, running-stat updates): library ops
        return self._embed_kernels(x)
