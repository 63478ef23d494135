"""TEST INFRASTRUCTURE ONLY (tests / smoke / bench cpu_baseline) - never imported by the product path.

CPU restatement of the 8 kHz -> 16 kHz resampling the reference applies to the reference utterance before RawNet3
(src/inferencers/inferencer_rawnet.py:21-23,36; src/trainers/trainer_rawnet.py:14-16,31:
``torchaudio.transforms.Resample(sample_rate, 16000, dtype=torch.float32)``).  torchaudio is a third-party dependency
of the reference (requirements.txt), not vendored under /root/reference; its published algorithm ("sinc_interp_hann":
band-limited sinc interpolation with a Hann window, lowpass_filter_width = 6, rolloff = 0.99) is restated here:

  orig, new   = orig_freq / gcd, new_freq / gcd
  base        = min(orig, new) * rolloff
  width       = ceil(lowpass_filter_width * orig / base)
  for phase i in 0..new-1, tap k in 0..2*width+orig-1:
      t       = clamp((-i / new + (k - width) / orig) * base, -lpw, +lpw)
      kernel[i,k] = sinc(pi t) * cos(pi t / (2 lpw))^2 * base / orig
  out[q*new + i] = sum_k kernel[i,k] * x_padded[q*orig + k],   x padded by `width` zeros left, `width + orig` right
  out length  = ceil(new * T / orig)

Pinned: tests/golden/resample_8k_16k.npz holds outputs of the installed torchaudio (2.11) on seeded inputs
(tests/golden/make_golden_resample.py); tests/test_resample_cpu.py checks this restatement against them.
"""
import math

import numpy as np


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """-> (kernel [new, taps] float32, width, orig, new).  Every step in float32, as Resample(dtype=float32) does."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = np.float32(min(orig, new) * rolloff)
    width = int(math.ceil(lowpass_filter_width * orig / float(base)))
    idx = np.arange(-width, width + orig, dtype=np.float32) / np.float32(orig)
    phase = np.arange(0, -new, -1, dtype=np.float32) / np.float32(new)
    t = (phase[:, None] + idx[None, :]) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width).astype(np.float32)
    window = np.cos(t * np.float32(math.pi) / np.float32(lowpass_filter_width) / np.float32(2)) ** 2
    tp = t * np.float32(math.pi)
    scale = np.float32(float(base) / orig)
    with np.errstate(invalid='ignore', divide='ignore'):
        k = np.where(tp == 0, np.float32(1.0), np.sin(tp) / tp).astype(np.float32)
    return (k * window * scale).astype(np.float32), width, orig, new


def resample(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """x [..., T] -> [..., ceil(new * T / orig)] (float64 accumulation of float32 taps)."""
    kern, width, orig, new = sinc_resample_kernel(orig_freq, new_freq)
    taps = kern.shape[1]
    x2 = np.asarray(x, dtype=np.float32).reshape(-1, x.shape[-1])
    T = x2.shape[1]
    xp = np.pad(x2, ((0, 0), (width, width + orig))).astype(np.float64)
    n_q = (xp.shape[1] - taps) // orig + 1
    out = np.zeros((x2.shape[0], n_q, new), dtype=np.float64)
    for k in range(taps):
        out += xp[:, k:k + (n_q - 1) * orig + 1:orig, None] * kern[None, None, :, k].astype(np.float64)
    out = out.reshape(x2.shape[0], -1)[:, :-(-new * T // orig)]
    return out.reshape(x.shape[:-1] + out.shape[-1:]).astype(np.float32)
