"""DPRNN-RawNet (cfg 4) on the CPU: the oracle's RawNet3 restatement against the fixture produced by the reference's
own RawNet3 / DPRNNRawNetTasNet classes (tests/golden/make_golden_rawnet.py), and the drop-in class's state_dict layout
and seeded weights.  The sinc filterbank itself is restated third-party code (parity unpinned, see the oracle header)."""
import json
import os
import sys

import torch

from conftest import GOLDEN, load_golden, weight_fingerprint
from oracle import dprnn_oracle as O
from oracle import rawnet_oracle as RO
import tss_with_dprnn_b200 as P

sys.path.insert(0, GOLDEN)
from rawnet_perturb import perturb_rawnet_state  # noqa: E402


def build(meta):
    torch.manual_seed(meta['wseed'])
    model = P.DPRNNRawNetTasNet(**meta['kwargs']).eval()
    fp0 = weight_fingerprint(model.state_dict())
    assert abs(fp0 - meta['weight_fingerprint_seeded']) <= 1e-9 * fp0, \
        'seeded weights differ from the reference constructor (module construction order changed?)'
    assert perturb_rawnet_state(model) == meta['perturbed']
    fp = weight_fingerprint(model.state_dict())
    assert abs(fp - meta['weight_fingerprint']) <= 1e-9 * fp
    return model


def test_rawnet_state_dict_layout_matches_reference():
    meta, _ = load_golden('rawnet_att_r1_eval')
    want = json.load(open(os.path.join(GOLDEN, 'state_dict_layout_rawnet.json')))
    torch.manual_seed(0)
    sd = P.DPRNNRawNetTasNet(**meta['kwargs']).state_dict()
    got = {k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}
    assert got == want


def test_rawnet_oracle_matches_reference_fixture():
    meta, arr = load_golden('rawnet_att_r1_eval')
    torch.set_num_threads(os.cpu_count())
    model = build(meta)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    kw = meta['kwargs']
    cfg = O.Config(n_repeats=kw['n_repeats'], fusion_type=kw['fusion_type'])
    mix, ref = torch.from_numpy(arr['mix']), torch.from_numpy(arr['ref'])
    with torch.no_grad():
        emb = RO.rawnet3_forward(ref, sd, 'separation.spk_encoder.')
        est, logits = RO.rawnet_tasnet_forward(mix, ref, sd, cfg)
    # log(|sinc conv| + 1e-6) is ill-conditioned near zero crossings: fp32 re-association noise reaches ~2e-5 here
    assert O.peak_rel_err(emb, torch.from_numpy(arr['emb'])) < 1e-4
    assert O.peak_rel_err(est, torch.from_numpy(arr['est'])) < 1e-4
    assert O.peak_rel_err(logits, torch.from_numpy(arr['logits'])) < 1e-4


def test_sinc_filterbank_properties():
    """Domain properties of the (unpinned) filterbank restatement: 128 even + 128 odd filters of 251 taps, unit
    pass-band normalisation at the centre tap, (anti)symmetry, mel-ordered bands."""
    fb = RO.ParamSincFB(256, 251, stride=10)
    f = fb.filters()[:, 0]
    assert f.shape == (256, 251)
    cos_f, sin_f = f[:128], f[128:]
    assert torch.allclose(cos_f, torch.flip(cos_f, dims=[1]))
    assert torch.allclose(sin_f, -torch.flip(sin_f, dims=[1]))
    assert torch.allclose(cos_f[:, 125], torch.ones(128)) and torch.all(sin_f[:, 125] == 0)
    low = 50 + fb.low_hz_.abs()[:, 0]
    assert torch.all(low[1:] > low[:-1])
