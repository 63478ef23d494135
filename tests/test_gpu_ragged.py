"""Ragged (variable-length) batches, SURVEY.md section 8d cfg 3: a packed batch of utterances of different lengths must
give, for every utterance, the result of the reference's B = 1 loop (src/inferencers/inferencer_spe.py:25-45).
Checked (a) against the CPU oracle per utterance and (b) bit-exactly against our own uniform B = 1 path - packing
must not change a single bit, in the exact-fp32 mode and in the bf16 tensor-core mode."""
import pytest
import torch

from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P

pytestmark = pytest.mark.gpu
KW = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
          n_repeats=2, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0)
LENS = [(3000, 2500), (4571, 3000), (251, 400), (8000, 8000), (2999, 1234)]


def waves(lens, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [0.05 * torch.randn(t, generator=g) for t in lens]


@pytest.mark.parametrize('fusion', ['cat', 'add', 'mul', 'film', 'att'])
@pytest.mark.parametrize('precision', ['fp32', 'bf16', 'fp16'])
def test_spe_ragged_equals_per_utterance(fusion, precision):
    torch.manual_seed(3)
    model = P.DPRNNSpeTasNet(**KW, fusion_type=fusion).eval().cuda()
    model.precision = precision
    mixes, refs = waves([a for a, _ in LENS], 1), waves([b for _, b in LENS], 2)
    with torch.no_grad():
        est, logits = model.forward_ragged([m.cuda() for m in mixes], [r.cuda() for r in refs])
        for b, (m, r) in enumerate(zip(mixes, refs)):
            e1, l1 = model(m[None].cuda(), r[None].cuda(), torch.tensor(float(r.numel())))
            assert est[b].shape == m.shape
            assert torch.equal(est[b], e1[0]), (b, float((est[b] - e1[0]).abs().max()))
            assert torch.equal(logits[b], l1[0])


def test_spe_ragged_matches_oracle():
    torch.manual_seed(5)
    kw = dict(KW, n_repeats=1)
    model = P.DPRNNSpeTasNet(**kw, fusion_type='att').eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    lens = [(1500, 900), (2777, 2000), (600, 3000)]
    mixes, refs = waves([a for a, _ in lens], 3), waves([b for _, b in lens], 4)
    with torch.no_grad():
        est, logits = model.forward_ragged([m.cuda() for m in mixes], [r.cuda() for r in refs])
        cfg = O.Config(n_repeats=1, fusion_type='att')
        for b, (m, r) in enumerate(zip(mixes, refs)):
            want, wl = O.spe_forward(m[None], r[None], torch.tensor(float(r.numel())), sd, cfg)
            assert O.peak_rel_err(est[b].cpu(), want[0]) < 2e-5
            assert O.peak_rel_err(logits[b].cpu(), wl[0]) < 2e-5


@pytest.mark.parametrize('precision', ['fp32', 'bf16', 'fp16'])
def test_ira_ragged_equals_per_utterance(precision):
    torch.manual_seed(6)
    model = P.DPRNNSpeIRATasNet(**dict(KW, n_repeats=1), fusion_type='cat').eval().cuda()
    model.precision = precision
    mixes, refs = waves([a for a, _ in LENS[:4]], 5), waves([b for _, b in LENS[:4]], 6)
    with torch.no_grad():
        est, logits = model.forward_ragged([m.cuda() for m in mixes], [r.cuda() for r in refs])
        for b, (m, r) in enumerate(zip(mixes, refs)):
            e1, l1 = model(m[None].cuda(), r[None].cuda(), torch.tensor(float(r.numel())))
            assert torch.equal(est[b], e1[0]), (b, float((est[b] - e1[0]).abs().max()))
            assert torch.equal(logits[b], l1[0])


@pytest.mark.parametrize('precision', ['fp32', 'bf16', 'fp16'])
def test_tasnet_ragged_equals_per_utterance(precision):
    torch.manual_seed(7)
    model = P.DPRNNTasNet(**dict(KW, n_repeats=1)).eval().cuda()
    model.precision = precision
    mixes = waves([2000, 5000, 333], 8)
    with torch.no_grad():
        est = model.forward_ragged([m.cuda() for m in mixes])
        for b, m in enumerate(mixes):
            e1 = model(m[None].cuda())
            assert est[b].shape == (2, m.numel())
            assert torch.equal(est[b], e1[0])


def test_ragged_rejects_train_mode_and_stride():
    model = P.DPRNNSpeTasNet(**KW, fusion_type='cat').train().cuda()
    with pytest.raises(NotImplementedError):
        model.forward_ragged([torch.zeros(1000).cuda()], [torch.zeros(1000).cuda()])


@pytest.mark.parametrize('precision', ['bf16', 'fp16'])
def test_ragged_fold_fused_is_bit_identical(precision):
    """Engine.fold_fused on packed batches (dprnn_unfold_ragged_h16 + dprnn_norm_residual_fold_prelu_ragged_h16): no fp32
    chunk-space tensor, no bit changed against unfold + cast ... norm (fp32 out) + fold."""
    torch.manual_seed(8)
    model = P.DPRNNSpeTasNet(**KW, fusion_type='film').eval().cuda()
    model.precision = precision
    mixes, refs = waves([a for a, _ in LENS], 7), waves([b for _, b in LENS], 8)
    outs = []
    with torch.no_grad():
        for fused in (True, False):
            model._engine.fold_fused = fused
            outs.append(model.forward_ragged([m.cuda() for m in mixes], [r.cuda() for r in refs]))
    (e1, l1), (e2, l2) = outs
    assert all(torch.equal(a, b) for a, b in zip(e1, e2)) and all(torch.equal(a, b) for a, b in zip(l1, l2))
