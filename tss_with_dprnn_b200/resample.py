"""The sample-rate conversion the RawNet inferencer / trainer apply to the reference utterance before the speaker
encoder (src/inferencers/inferencer_rawnet.py:21-23,36; src/trainers/trainer_rawnet.py:14-16,31 build
``torchaudio.transforms.Resample(sample_rate, 16000, dtype=torch.float32)`` and call it on the CPU): same constructor
arguments and call signature, the filtering runs as one CUDA kernel so that cfg 4 stays on the device end to end.

The windowed-sinc taps follow torchaudio's published 'sinc_interp_hann' recipe (band-limited interpolation, Hann window,
lowpass_filter_width 6, rolloff 0.99), computed once on the host in float32 like Resample(dtype=float32) does.
"""
from __future__ import annotations

import math

import torch

from ._lib import lib


def _sinc_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int, rolloff: float):
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    if lowpass_filter_width <= 0:
        raise ValueError('Low pass filter width should be positive.')
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    f32 = torch.float32
    idx = torch.arange(-width, width + orig, dtype=f32)[None] / orig
    t = torch.arange(0, -new, -1, dtype=f32)[:, None] / new + idx
    t = (t * base).clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    kern = torch.where(t == 0, torch.tensor(1.0, dtype=f32), t.sin() / t)
    return (kern * window * (base / orig)).contiguous(), width, orig, new


class Resample(torch.nn.Module):
    """Drop-in for ``torchaudio.transforms.Resample(orig_freq, new_freq, dtype=torch.float32)`` on CUDA tensors."""

    def __init__(self, orig_freq: int = 16000, new_freq: int = 16000, resampling_method: str = 'sinc_interp_hann',
                 lowpass_filter_width: int = 6, rolloff: float = 0.99, beta=None, *, dtype=None):
        super().__init__()
        if resampling_method not in ('sinc_interp_hann', 'sinc_interpolation'):
            raise NotImplementedError("only the reference's default resampling_method ('sinc_interp_hann') is built")
        if dtype not in (None, torch.float32):
            raise NotImplementedError('the kernel is built for float32 taps (what the reference constructs)')
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        kern, self.width, self.orig, self.new = _sinc_kernel(orig_freq, new_freq, lowpass_filter_width, rolloff)
        self.register_buffer('kernel', kern, persistent=False)

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        if self.orig_freq == self.new_freq:
            return waveform
        if not waveform.is_floating_point():
            raise TypeError(f'Expected floating point type for waveform tensor, but received {waveform.dtype}.')
        if not waveform.is_cuda:
            raise RuntimeError('tss_with_dprnn_b200.Resample runs on CUDA tensors only (no CPU path)')
        shape = waveform.shape
        x = waveform.reshape(-1, shape[-1]).contiguous().float()
        B, T = x.shape
        To = -(-self.new * T // self.orig)
        out = torch.empty((B, To), device=x.device, dtype=torch.float32)
        kern = self.kernel.to(x.device)
        lib().call('dprnn_resample_fir', x, kern, out, B, T, To, self.orig, self.new, kern.shape[1], self.width,
                   torch.cuda.current_stream().cuda_stream)
        return out.view(shape[:-1] + (To,))
