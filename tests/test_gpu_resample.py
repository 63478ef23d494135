"""The CUDA resampler against the torchaudio golden outputs and the oracle restatement."""
import os

import numpy as np
import pytest
import torch

import tss_with_dprnn_b200 as P
from oracle.resample_oracle import resample

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'resample_8k_16k.npz'))


@pytest.mark.parametrize('case', list('abcd'))
def test_resample_matches_golden(case):
    o, n = (int(v) for v in Z[case + '_rates'])
    r = P.Resample(o, n, dtype=torch.float32).cuda()
    y = r(torch.from_numpy(Z[case + '_x']).cuda()).cpu().numpy()
    assert y.shape == Z[case + '_y'].shape
    assert np.abs(y - Z[case + '_y']).max() < 1e-6 * np.abs(Z[case + '_y']).max()


def test_resample_full_size_and_batch_shapes():
    """cfg-4 size: 16 references of 3 s @ 8 kHz -> 48000 samples @ 16 kHz, leading dims preserved."""
    g = torch.Generator().manual_seed(5)
    x = 0.05 * torch.randn(4, 4, 24000, generator=g)
    y = P.Resample(8000, 16000, dtype=torch.float32)(x.cuda())
    assert y.shape == (4, 4, 48000)
    want = resample(x.numpy(), 8000, 16000)
    assert np.abs(y.cpu().numpy() - want).max() < 1e-6 * np.abs(want).max()
    with pytest.raises(RuntimeError):
        P.Resample(8000, 16000)(x)                       # CPU tensor: no CPU path
    assert P.Resample(8000, 8000)(x) is x
