"""Constructor options beyond the shipped YAMLs (chunk length / hop, encoder width, unidirectional inter-chunk RNN,
gLN + ReLU): exact-fp32 mode against the CPU oracle, bf16 mode close to it, and a packed ragged batch bit-identical to
the per-utterance forward - for the 'cat' and the 'att' fusion."""
import pytest
import torch

from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P

pytestmark = pytest.mark.gpu
BASE = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
            n_repeats=1, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0)
CASES = [dict(chunk_length=200, hop_length=100), dict(chunk_length=100, hop_length=50), dict(input_size=128),
         dict(input_size=32), dict(bidirectional=False), dict(norm_type='gLN', activation_type='relu'),
         dict(chunk_length=256, hop_length=128), dict(chunk_length=64, hop_length=32)]


@pytest.mark.parametrize('case', range(len(CASES)))
@pytest.mark.parametrize('fusion', ['cat', 'att'])
def test_config(case, fusion):
    kw = dict(BASE, **CASES[case])
    torch.manual_seed(case)
    m = P.DPRNNSpeTasNet(**kw, fusion_type=fusion).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    g = torch.Generator().manual_seed(1)
    mix, ref = 0.05 * torch.randn(3, 5001, generator=g), 0.05 * torch.randn(3, 4000, generator=g)
    cfg = O.Config(**{k: kw[k] for k in ('input_size', 'feature_size', 'hidden_size', 'chunk_length', 'kernel_size',
                                         'hop_length', 'n_repeats', 'bidirectional', 'norm_type', 'activation_type')},
                   fusion_type=fusion)
    with torch.no_grad():
        want, wl = O.spe_forward(mix, ref, torch.tensor(4000.), sd, cfg)
        m.precision = 'fp32'
        est, lg = m(mix.cuda(), ref.cuda(), torch.tensor(4000.))
        assert O.peak_rel_err(est.cpu(), want) < 2e-5
        assert O.peak_rel_err(lg.cpu(), wl) < 2e-5
        m.precision = 'bf16'
        est, _ = m(mix.cuda(), ref.cuda(), torch.tensor(4000.))
        assert O.peak_rel_err(est.cpu(), want) < 2e-2
        rag, _ = m.forward_ragged([mix[0].cuda(), mix[1, :3000].cuda()], [ref[0].cuda(), ref[1, :2500].cuda()])
        e1, _ = m(mix[1:2, :3000].cuda(), ref[1:2, :2500].cuda(), torch.tensor(2500.))
        assert torch.equal(rag[1], e1[0])
