"""The callers' side of the path: GPU SI-SDR, the batched evaluation driver, and the fused
clip + Adam step, against the oracle / torch reference implementations."""
import pytest
import torch

from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P
from tss_with_dprnn_b200.evaluate import evaluate, si_sdr
from tss_with_dprnn_b200.dp import ClipAdam, FlatParams

pytestmark = pytest.mark.gpu
KW = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
          n_repeats=1, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0)


def test_si_sdr_matches_formula():
    g = torch.Generator().manual_seed(0)
    t = torch.randn(5, 24000, generator=g)
    e = t * torch.tensor([1.0, 0.5, 2.0, 1.0, -1.0]).view(-1, 1) + torch.tensor([0.01, 0.1, 0.3, 1.0, 0.2]).view(-1, 1) * \
        torch.randn(5, 24000, generator=g) + 0.3
    want = O.si_sdr_db(e, t)
    got = si_sdr(e.cuda(), t.cuda()).cpu()
    assert torch.allclose(got.double(), want, atol=1e-3)
    lens = [24000, 1000, 17, 5000]
    ee, tt = [e[i, :n] for i, n in enumerate(lens)], [t[i, :n] for i, n in enumerate(lens)]
    got = si_sdr(torch.cat(ee).cuda(), torch.cat(tt).cuda(), lens).cpu()
    want = torch.stack([O.si_sdr_db(a[None], b[None])[0] for a, b in zip(ee, tt)])
    assert torch.allclose(got.double(), want, atol=1e-3)


def test_evaluate_driver_equals_per_utterance_loop():
    torch.manual_seed(1)
    model = P.DPRNNSpeTasNet(**KW, fusion_type='film').eval().cuda()
    g = torch.Generator().manual_seed(2)
    lens = [(3000, 2500), (4571, 3000), (800, 900), (6000, 6000), (2999, 1234), (3500, 3500), (1200, 4000)]
    mixes = [0.05 * torch.randn(a, generator=g) for a, _ in lens]
    refs = [0.05 * torch.randn(b, generator=g) for _, b in lens]
    tgts = [m + 0.01 * torch.randn(m.shape, generator=g) for m in mixes]
    res = {}
    for rank in range(2):          # two ranks of a sharded evaluation, run one after the other
        for r in evaluate(model, mixes, refs, tgts, bucket=3, rank=rank, world=2, keep_audio=True):
            assert r['index'] not in res
            res[r['index']] = r
    assert sorted(res) == list(range(len(lens)))
    with torch.no_grad():
        for i, (m, r, t) in enumerate(zip(mixes, refs, tgts)):
            est, logits = model(m[None].cuda(), r[None].cuda(), torch.tensor(float(r.numel())))
            assert torch.equal(res[i]['estimate'], est[0].cpu())
            assert torch.equal(res[i]['logits'], logits[0].cpu())
            assert abs(res[i]['si_sdr'] - float(O.si_sdr_db(est.cpu(), t[None])[0])) < 1e-3


def test_evaluate_driver_bss_pit():
    torch.manual_seed(3)
    model = P.DPRNNTasNet(**KW).eval().cuda()
    g = torch.Generator().manual_seed(4)
    mixes = [0.05 * torch.randn(n, generator=g) for n in (2000, 3100, 900)]
    with torch.no_grad():
        own = [model(m[None].cuda())[0].cpu() for m in mixes]
    tgts = [torch.stack([o[1], o[0]]) for o in own]            # swapped speakers: PIT must find the permutation
    res = evaluate(model, mixes, None, tgts, bucket=2)
    assert all(r['si_sdr'] > 60 for r in res)


def test_clip_adam_matches_torch():
    torch.manual_seed(5)
    ref = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.PReLU(), torch.nn.Linear(128, 251)).cuda()
    ours = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.PReLU(), torch.nn.Linear(128, 251)).cuda()
    ours.load_state_dict(ref.state_dict())
    opt = torch.optim.Adam(ref.parameters(), lr=5e-4, weight_decay=1e-5)
    fp = FlatParams(ours)
    mine = ClipAdam(fp, lr=5e-4, weight_decay=1e-5, max_norm=5.0)
    g = torch.Generator(device='cuda').manual_seed(6)
    for it in range(5):
        scale = [30.0, 0.1, 5.0, 100.0, 1.0][it]              # clipped and unclipped iterations
        grads = [scale * torch.randn(p.shape, generator=g, device='cuda') for p in ref.parameters()]
        for p, gr in zip(ref.parameters(), grads):
            p.grad = gr.clone()
        for (_, p), gr in zip(fp.named, grads):
            p.grad.copy_(gr)
        tn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 5.0)
        opt.step()
        got = mine.step()
        assert torch.allclose(got[0], tn, rtol=1e-5)
        for a, b in zip(ref.parameters(), ours.parameters()):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), it
