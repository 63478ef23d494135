"""CPU oracle (TEST INFRASTRUCTURE - never imported by the product path) for the RawNet3 speaker encoder of
DPRNN-RawNet (cfg 4).

Two parts:

* ``ParamSincFB`` / ``Encoder``: a restatement of the two classes the reference imports from the THIRD-PARTY package
  ``asteroid_filterbanks==0.4.0`` (requirements.txt:2; src/models/rawnet/RawNet3.py:5,26-32), which is neither vendored
  under /root/reference nor installed here.  **PARITY UNPINNED** for this part: it follows the published definition of
  the parameterised sinc filterbank (SincNet band-pass filters on a mel-initialised grid, learnable low cut-off and
  bandwidth, Hamming half-window, an even "cos" and an odd "sin" filter per band; kernel 251, stride 10), and nothing in
  the reference (no test, golden vector, checkpoint or metric) constrains its arithmetic here.
* ``rawnet3_forward`` / ``rawnet_tasnet_forward``: a functional restatement of RawNet3.forward
  (src/models/rawnet/RawNet3.py:72-136), Bottle2neck / AFMS / PreEmphasis (RawNetBasicBlock.py:8-142) and
  DPRNNRawNetTasNet.forward (src/models/dprnn_rawnet.py:72-105,171-182).  PINNED: tests/golden/make_golden_rawnet.py
  imports the reference's own RawNet3 / DPRNNRawNetTasNet classes with ``asteroid_filterbanks`` stubbed by the two
  classes above and commits their outputs; tests/test_oracle_vs_golden.py holds this file to those fixtures.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import dprnn_oracle as O


# ---------------------------------------------------------------------------------------------------------
# asteroid_filterbanks restatement (parity unpinned)
# ---------------------------------------------------------------------------------------------------------
class ParamSincFB(nn.Module):
    """Parameterised sinc filterbank: n_filters/2 band-pass pairs (cos = even, sin = odd)."""

    def __init__(self, n_filters, kernel_size, stride=None, sample_rate=16000.0, min_low_hz=50, min_band_hz=50):
        super().__init__()
        if kernel_size % 2 == 0:
            kernel_size += 1
        self.n_filters, self.kernel_size = n_filters, kernel_size
        self.stride = stride if stride else kernel_size // 2
        self.sample_rate = sample_rate
        self.min_low_hz, self.min_band_hz = min_low_hz, min_band_hz
        self.half_kernel = kernel_size // 2
        self.cutoff = n_filters // 2
        low_hz = 30.0
        high_hz = sample_rate / 2 - (min_low_hz + min_band_hz)
        mel = np.linspace(self.to_mel(low_hz), self.to_mel(high_hz), n_filters // 2 + 1, dtype='float32')
        hz = self.to_hz(mel)
        self.low_hz_ = nn.Parameter(torch.from_numpy(hz[:-1]).view(-1, 1))
        self.band_hz_ = nn.Parameter(torch.from_numpy(np.diff(hz)).view(-1, 1))
        window_ = np.hamming(kernel_size)[: self.half_kernel]                 # half window
        n_ = 2 * np.pi * (torch.arange(-self.half_kernel, 0.0).view(1, -1) / sample_rate)   # half time axis
        self.register_buffer('window_', torch.from_numpy(window_).float())
        self.register_buffer('n_', n_)

    @staticmethod
    def to_mel(hz):
        return 2595 * np.log10(1 + hz / 700)

    @staticmethod
    def to_hz(mel):
        return 700 * (10 ** (mel / 2595) - 1)

    def make_filters(self, low, high, filt_type):
        band = (high - low)[:, 0]
        ft_low = torch.matmul(low, self.n_)
        ft_high = torch.matmul(high, self.n_)
        if filt_type == 'cos':
            bp_left = ((torch.sin(ft_high) - torch.sin(ft_low)) / (self.n_ / 2)) * self.window_
            bp_center = 2 * band.view(-1, 1)
            bp_right = torch.flip(bp_left, dims=[1])
        else:
            bp_left = ((torch.cos(ft_low) - torch.cos(ft_high)) / (self.n_ / 2)) * self.window_
            bp_center = torch.zeros_like(band.view(-1, 1))
            bp_right = -torch.flip(bp_left, dims=[1])
        band_pass = torch.cat([bp_left, bp_center, bp_right], dim=1)
        band_pass = band_pass / (2 * band[:, None])
        return band_pass.view(self.n_filters // 2, 1, self.kernel_size)

    def filters(self):
        low = self.min_low_hz + torch.abs(self.low_hz_)
        high = torch.clamp(low + self.min_band_hz + torch.abs(self.band_hz_), self.min_low_hz, self.sample_rate / 2)
        return torch.cat([self.make_filters(low, high, 'cos'), self.make_filters(low, high, 'sin')], dim=0)


class Encoder(nn.Module):
    """asteroid_filterbanks.Encoder: strided conv1d of the waveform with the filterbank's filters, no padding."""

    def __init__(self, filterbank):
        super().__init__()
        self.filterbank = filterbank

    def forward(self, waveform):
        if waveform.dim() == 2:
            waveform = waveform.unsqueeze(1)
        return F.conv1d(waveform, self.filterbank.filters(), stride=self.filterbank.stride)


def sinc_filters(sd, prefix, sample_rate=16000.0, min_low_hz=50, min_band_hz=50):
    """filters [n_filters, 1, kernel] from the state_dict entries of conv1.filterbank (functional form of the above)."""
    low_hz_, band_hz_ = sd[prefix + 'low_hz_'], sd[prefix + 'band_hz_']
    window_, n_ = sd[prefix + 'window_'], sd[prefix + 'n_']
    low = min_low_hz + torch.abs(low_hz_)
    high = torch.clamp(low + min_band_hz + torch.abs(band_hz_), min_low_hz, sample_rate / 2)
    band = (high - low)[:, 0]
    ft_low, ft_high = torch.matmul(low, n_), torch.matmul(high, n_)
    cos_left = ((torch.sin(ft_high) - torch.sin(ft_low)) / (n_ / 2)) * window_
    sin_left = ((torch.cos(ft_low) - torch.cos(ft_high)) / (n_ / 2)) * window_
    cos_f = torch.cat([cos_left, 2 * band.view(-1, 1), torch.flip(cos_left, dims=[1])], dim=1) / (2 * band[:, None])
    sin_f = torch.cat([sin_left, torch.zeros_like(band.view(-1, 1)), -torch.flip(sin_left, dims=[1])], dim=1) / (2 * band[:, None])
    return torch.cat([cos_f, sin_f], dim=0).unsqueeze(1)


# ---------------------------------------------------------------------------------------------------------
# RawNet3 (pinned against the reference's own classes)
# ---------------------------------------------------------------------------------------------------------
def _bn(x, sd, p, eps=1e-5):
    """nn.BatchNorm1d in eval mode on [B,C,T] or [B,C]."""
    shape = (1, -1, 1) if x.dim() == 3 else (1, -1)
    return (x - sd[p + 'running_mean'].view(shape)) / torch.sqrt(sd[p + 'running_var'].view(shape) + eps) \
        * sd[p + 'weight'].view(shape) + sd[p + 'bias'].view(shape)


def bottle2neck(x, sd, p, dilation, pool, scale=8):
    """Bottle2neck.forward, RawNetBasicBlock.py:111-142 (Res2Net block, kernel 3) + AFMS (:48-55)."""
    planes = sd[p + 'conv3.weight'].shape[0]
    width = planes // scale
    residual = F.conv1d(x, sd[p + 'residual.0.weight']) if (p + 'residual.0.weight') in sd else x
    out = _bn(torch.relu(F.conv1d(x, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'])), sd, p + 'bn1.')
    spx = torch.split(out, width, 1)
    outs = []
    sp = None
    for i in range(scale - 1):
        sp = spx[i] if i == 0 else sp + spx[i]
        sp = F.conv1d(sp, sd[p + f'convs.{i}.weight'], sd[p + f'convs.{i}.bias'], dilation=dilation, padding=dilation)
        sp = _bn(torch.relu(sp), sd, p + f'bns.{i}.')
        outs.append(sp)
    outs.append(spx[scale - 1])
    out = torch.cat(outs, 1)
    out = _bn(torch.relu(F.conv1d(out, sd[p + 'conv3.weight'], sd[p + 'conv3.bias'])), sd, p + 'bn3.')
    out = out + residual
    if pool:
        out = F.max_pool1d(out, pool)
    y = torch.sigmoid(F.linear(out.mean(-1), sd[p + 'afms.fc.weight'], sd[p + 'afms.fc.bias']))
    return (out + sd[p + 'afms.alpha']) * y.unsqueeze(-1)


def rawnet3_forward(x, sd, p):
    """RawNet3.forward (RawNet3.py:72-136) in eval mode with the DPRNNRawNet settings (dprnn_rawnet.py:57-70):
    context=True, summed=True, encoder_type='ECA', out_bn=False, log_sinc=True, norm_sinc='mean', sinc_stride=10.
    x [B, T] raw 16 kHz reference -> [B, nOut]."""
    # PreEmphasis (RawNetBasicBlock.py:20-28): reflect-pad one sample on the left, y[t] = x[t] - 0.97 x[t-1]
    xi = F.pad(x.unsqueeze(1), (1, 0), 'reflect')
    xi = F.conv1d(xi, sd[p + 'preprocess.0.flipped_filter'])
    # InstanceNorm1d(1, eps=1e-4, affine=True) (RawNet3.py:23-25)
    m, v = xi.mean(-1, keepdim=True), xi.var(-1, unbiased=False, keepdim=True)
    xi = (xi - m) / torch.sqrt(v + 1e-4) * sd[p + 'preprocess.1.weight'].view(1, -1, 1) + sd[p + 'preprocess.1.bias'].view(1, -1, 1)
    f = torch.abs(F.conv1d(xi, sinc_filters(sd, p + 'conv1.filterbank.'), stride=10))      # :79
    f = torch.log(f + 1e-6)                                                              # :81
    f = f - f.mean(-1, keepdim=True)                                                     # :83
    x1 = bottle2neck(f, sd, p + 'layer1.', 2, 5)
    x2 = bottle2neck(x1, sd, p + 'layer2.', 3, 3)
    x3 = bottle2neck(F.max_pool1d(x1, 3) + x2, sd, p + 'layer3.', 4, 0)                   # summed (:93)
    h = torch.relu(F.conv1d(torch.cat((F.max_pool1d(x1, 3), x2, x3), 1), sd[p + 'layer4.weight'], sd[p + 'layer4.bias']))
    t = h.shape[-1]
    g = torch.cat((h, h.mean(2, keepdim=True).repeat(1, 1, t),
                   torch.sqrt(h.var(2, keepdim=True).clamp(min=1e-4, max=1e4)).repeat(1, 1, t)), 1)     # :105-117
    w = F.conv1d(g, sd[p + 'attention.0.weight'], sd[p + 'attention.0.bias'])
    w = _bn(torch.relu(w), sd, p + 'attention.2.')
    w = torch.softmax(F.conv1d(w, sd[p + 'attention.3.weight'], sd[p + 'attention.3.bias']), dim=2)
    mu = torch.sum(h * w, 2)
    sg = torch.sqrt((torch.sum(h ** 2 * w, 2) - mu ** 2).clamp(min=1e-4, max=1e4))
    e = _bn(torch.cat((mu, sg), 1), sd, p + 'bn5.')
    return F.linear(e, sd[p + 'fc6.weight'], sd[p + 'fc6.bias'])                         # out_bn=False: bn6 unused


def rawnet_tasnet_forward(mix, ref16k, sd, cfg: O.Config):
    """DPRNNRawNetTasNet.forward (dprnn_rawnet.py:171-182): masker of DPRNNSpe with aux = RawNet3(raw reference)."""
    emb = rawnet3_forward(ref16k, sd, 'separation.spk_encoder.')
    return O.spe_forward(mix, None, None, sd, cfg, embedding=emb)
