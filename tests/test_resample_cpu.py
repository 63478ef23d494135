"""The resampler restatement (oracle/resample_oracle.py) and the product's host-side tap computation against golden
outputs of the installed torchaudio (tests/golden/make_golden_resample.py)."""
import os

import numpy as np
import pytest

from oracle.resample_oracle import resample, sinc_resample_kernel

Z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'resample_8k_16k.npz'))


@pytest.mark.parametrize('case', list('abcd'))
def test_oracle_matches_torchaudio_golden(case):
    o, n = (int(v) for v in Z[case + '_rates'])
    kern, width, orig, new = sinc_resample_kernel(o, n)
    assert kern.shape == Z[case + '_kernel'].shape
    assert np.abs(kern - Z[case + '_kernel']).max() < 1e-7
    y = resample(Z[case + '_x'], o, n)
    assert y.shape == Z[case + '_y'].shape                          # ceil(new * T / orig)
    assert np.abs(y - Z[case + '_y']).max() < 1e-6 * np.abs(Z[case + '_y']).max()


@pytest.mark.parametrize('case', list('abcd'))
def test_product_taps_match_golden(case):
    from tss_with_dprnn_b200.resample import Resample
    o, n = (int(v) for v in Z[case + '_rates'])
    r = Resample(o, n, dtype=__import__('torch').float32)
    assert tuple(r.kernel.shape) == Z[case + '_kernel'].shape
    assert np.abs(r.kernel.numpy() - Z[case + '_kernel']).max() < 1e-7
