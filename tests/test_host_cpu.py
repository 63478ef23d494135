"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol the header
declares, the model classes expose the reference's state_dict layout, and the product path refuses
to run without CUDA (no CPU fallback)."""
import ctypes
import json
import os

import pytest
import torch

from conftest import GOLDEN, ROOT
import tss_with_dprnn_b200 as P
from tss_with_dprnn_b200._lib import LIB_PATH, parse_header

KW = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
          n_repeats=6, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB_PATH), 'run `python -m tss_with_dprnn_b200.build` first'
    protos = parse_header()
    assert len(protos) >= 20
    cdll = ctypes.CDLL(LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), name
    info = P.lib().build_info()
    assert 'sm_100a' in info


def test_state_dict_layout_matches_reference():
    """tests/golden/state_dict_layout.json was enumerated from the live reference classes."""
    layout = json.load(open(os.path.join(GOLDEN, 'state_dict_layout.json')))
    builders = {
        'tasnet': lambda: P.DPRNNTasNet(**KW),
        'ira_cat': lambda: P.DPRNNSpeIRATasNet(**KW, fusion_type='cat'),
    }
    for ft in ('cat', 'add', 'mul', 'film', 'att'):
        builders[f'spe_{ft}'] = (lambda ft=ft: P.DPRNNSpeTasNet(**KW, fusion_type=ft))
        builders[f'spe_{ft}_gLN_uni'] = (lambda ft=ft: P.DPRNNSpeTasNet(**dict(KW, norm_type='gLN', bidirectional=False),
                                                                   fusion_type=ft))
    for name, want in layout.items():
        sd = builders[name]().state_dict()
        got = {k: list(v.shape) for k, v in sd.items()}
        assert got == want, name


def test_constructor_contract():
    with pytest.raises(NotImplementedError):
        P.DPRNNTasNet(**dict(KW, rnn_type='GRU'))
    with pytest.raises(ValueError):
        P.DPRNNSpeTasNet(**KW, fusion_type='nope')
    m = P.DPRNNTasNet(input_size=64, chunk_length=250)
    assert m.cfg['hop_length'] == 125 and m.stride == 1        # defaults: dprnn.py:127,243
    att = P.DPRNNSpeTasNet(**KW, fusion_type='att')
    assert not att.separation.average.weight.requires_grad     # frozen averaging conv, dprnn_spe.py:102-104
    assert torch.equal(att.separation.average.weight, torch.full((64, 1, 2), 0.5))


def test_no_cpu_path():
    m = P.DPRNNSpeTasNet(**dict(KW, n_repeats=1)).eval()
    with pytest.raises(RuntimeError, match='no CPU path'):
        m(torch.zeros(1, 4000), torch.zeros(1, 4000), torch.tensor(4000.))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'tss_with_dprnn_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in text.replace('no oracle', ''), f
