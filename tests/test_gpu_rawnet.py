"""DPRNN-RawNet (cfg 4) on the GPU: DPRNNRawNetTasNet.forward(mix, ref16k) against the fixture produced by the
reference's own classes (sinc filterbank stubbed by the restatement - parity unpinned for that part only)."""
import sys

import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P

sys.path.insert(0, GOLDEN)
from test_rawnet_cpu import build  # noqa: E402

pytestmark = pytest.mark.gpu


def test_rawnet_matches_reference_fixture_fp32():
    meta, arr = load_golden('rawnet_att_r1_eval')
    model = build(meta).cuda()
    mix, ref = torch.from_numpy(arr['mix']).cuda(), torch.from_numpy(arr['ref']).cuda()
    with torch.no_grad():
        n0 = P.lib().launches
        est, logits = model(mix, ref)
        assert P.lib().launches > n0            # the masker ran through the C ABI
        emb = model.separation.spk_encoder.embed(ref)
    assert O.peak_rel_err(emb.cpu(), torch.from_numpy(arr['emb'])) < 2e-4
    assert O.peak_rel_err(est.cpu(), torch.from_numpy(arr['est'])) < 2e-4
    assert O.peak_rel_err(logits.cpu(), torch.from_numpy(arr['logits'])) < 2e-4


def test_rawnet_bf16_mode_close():
    meta, arr = load_golden('rawnet_att_r1_eval')
    model = build(meta).cuda()
    model.precision = 'bf16'
    mix, ref = torch.from_numpy(arr['mix']).cuda(), torch.from_numpy(arr['ref']).cuda()
    with torch.no_grad():
        est, _ = model(mix, ref)
    want = torch.from_numpy(arr['est'])
    assert O.si_sdr_db(est.cpu(), want).min() > 30.0


def test_rawnet_rejects_cpu_and_train():
    meta, arr = load_golden('rawnet_att_r1_eval')
    torch.manual_seed(0)
    model = P.DPRNNRawNetTasNet(**meta['kwargs'])
    with pytest.raises(RuntimeError, match='no CPU path'):
        model.eval()(torch.zeros(1, 4000), torch.zeros(1, 8000))
    with pytest.raises(NotImplementedError):
        model.train().cuda()(torch.zeros(1, 4000).cuda(), torch.zeros(1, 8000).cuda())
