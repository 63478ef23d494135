"""RawNet3 speaker encoder of DPRNN-RawNet (src/models/rawnet/RawNet3.py, RawNetBasicBlock.py; cfg 4).

Parameter containers with the reference's module tree (so ``chkpts/dprnn-rawnet/*.pt`` strict-load and seeded
default initialisation reproduces the reference's weights), plus the forward.

The front-end (PreEmphasis, InstanceNorm, parameterised sinc filterbank, log-abs, mean normalisation) runs as
hand-written kernels (csrc/rawnet.cu).  STAGE 1 for the rest: RawNet3 is not
on north_star's kernel list (~8 % of the model's flops, run once per enrolment utterance); its Res2Net blocks and the
attentive statistics pooling are restated here with torch ops ON THE GPU (cuDNN / cuBLAS library calls) and produce the
[B, nOut] embedding - everything downstream of the embedding (attention fusion, the whole masker, decoder) is the
hand-written CUDA path.  Hand kernels for the Res2Net blocks are the first "next" row of section 8f.  There is no CPU
path: the tensors must be CUDA tensors.

The sinc front-end restates ``asteroid_filterbanks==0.4.0`` ``ParamSincFB`` / ``Encoder`` (third-party, not vendored in
the reference, not installable here): parity for that part is UNPINNED.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
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

    def run(self, x):
        bn = lambda t, m: F.batch_norm(t, m.running_mean, m.running_var, m.weight, m.bias, False, 0.0, m.eps)
        residual = x if isinstance(self.residual, nn.Identity) else F.conv1d(x, self.residual[0].weight)
        out = bn(torch.relu(F.conv1d(x, self.conv1.weight, self.conv1.bias)), self.bn1)
        spx = torch.split(out, self.width, 1)
        outs, sp = [], None
        for i in range(self.nums):
            sp = spx[i] if i == 0 else sp + spx[i]
            c = self.convs[i]
            sp = bn(torch.relu(F.conv1d(sp, c.weight, c.bias, dilation=self.dilation, padding=self.dilation)), self.bns[i])
            outs.append(sp)
        outs.append(spx[self.nums])
        out = bn(torch.relu(F.conv1d(torch.cat(outs, 1), self.conv3.weight, self.conv3.bias)), self.bn3) + residual
        if self.pool:
            out = F.max_pool1d(out, self.pool)
        y = torch.sigmoid(F.linear(out.mean(-1), self.afms.fc.weight, self.afms.fc.bias))
        return (out + self.afms.alpha) * y.unsqueeze(-1)


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
        self.allow_tf32 = False                     # library convolutions in full fp32 unless the model runs in bf16 mode

    @torch.no_grad()
    def embed(self, x):
        """RawNet3.forward (RawNet3.py:72-136), eval mode, on the GPU.  x [B, T] raw 16 kHz reference -> [B, nOut]."""
        if not x.is_cuda:
            raise RuntimeError('aux must be a CUDA tensor: tss_with_dprnn_b200 has no CPU path')
        if self.training:
            raise NotImplementedError('the RawNet3 speaker encoder runs in eval() mode (InferencerRawNet calls '
                                      'model.eval(), src/inferencers/inferencer_rawnet.py:29)')
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=self.allow_tf32):
            return self._embed(x)

    def _embed(self, x):
        bn = lambda t, m: F.batch_norm(t, m.running_mean, m.running_var, m.weight, m.bias, False, 0.0, m.eps)
        # front-end (PreEmphasis, InstanceNorm, sinc filterbank, log-abs, mean normalisation): hand-written kernels
        fb, inorm = self.conv1.filterbank, self.preprocess[1]
        x = x.contiguous()
        B, T = x.shape
        Tp = (T - fb.kernel_size) // fb.stride + 1
        feat = torch.empty((B, Tp, fb.n_filters), device=x.device, dtype=torch.float32)
        filt = torch.empty(fb.kernel_size * fb.n_filters, device=x.device, dtype=torch.float32)
        stats = torch.empty(2 * B, device=x.device, dtype=torch.float32)
        lib().call('dprnn_rawnet_frontend', x, B, T, inorm.weight.detach(), inorm.bias.detach(), fb.low_hz_.detach(),
                   fb.band_hz_.detach(), fb.window_, fb.n_, fb.n_filters, fb.kernel_size, fb.stride,
                   float(fb.sample_rate), filt, stats, feat, torch.cuda.current_stream().cuda_stream)
        f = feat.permute(0, 2, 1)                     # [B, 256, T'] view for the (stage-1) Res2Net blocks below
        x1 = self.layer1.run(f)
        x2 = self.layer2.run(x1)
        x1p = F.max_pool1d(x1, 3)
        x3 = self.layer3.run(x1p + x2)
        h = torch.relu(F.conv1d(torch.cat((x1p, x2, x3), 1), self.layer4.weight, self.layer4.bias))
        t = h.shape[-1]
        g = torch.cat((h, h.mean(2, keepdim=True).repeat(1, 1, t),
                       torch.sqrt(h.var(2, keepdim=True).clamp(min=1e-4, max=1e4)).repeat(1, 1, t)), 1)
        a = self.attention
        w = bn(torch.relu(F.conv1d(g, a[0].weight, a[0].bias)), a[2])
        w = torch.softmax(F.conv1d(w, a[3].weight, a[3].bias), dim=2)
        mu = torch.sum(h * w, 2)
        sg = torch.sqrt((torch.sum(h ** 2 * w, 2) - mu ** 2).clamp(min=1e-4, max=1e4))
        e = bn(torch.cat((mu, sg), 1), self.bn5)
        return F.linear(e, self.fc6.weight, self.fc6.bias).contiguous()
