"""Golden vectors for the 8 kHz -> 16 kHz resampler: outputs of the INSTALLED torchaudio
(transforms.Resample(8000, 16000, dtype=float32), as src/inferencers/inferencer_rawnet.py:21-23 builds it) on seeded inputs.
Run in the build container:  python tests/golden/make_golden_resample.py"""
import os

import numpy as np
import torch
import torchaudio.transforms as T

HERE = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(2024)
out = {}
for name, (o, n, shape) in {'a': (8000, 16000, (3, 4001)), 'b': (8000, 16000, (1, 17)), 'c': (16000, 8000, (2, 1000)),
                            'd': (8000, 12000, (2, 501))}.items():
    x = 0.1 * torch.randn(*shape, generator=g)
    r = T.Resample(o, n, dtype=torch.float32)
    out[name + '_x'] = x.numpy()
    out[name + '_y'] = r(x).numpy()
    out[name + '_kernel'] = r.kernel.numpy()[:, 0]
    out[name + '_rates'] = np.array([o, n])
np.savez_compressed(os.path.join(HERE, 'resample_8k_16k.npz'), **out)
print({k: v.shape for k, v in out.items()})
