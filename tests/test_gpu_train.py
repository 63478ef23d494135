"""cfg 5: the hand-written backward of DPRNN-Spe against autograd through the CPU oracle (fp64), parameter by parameter,
and one optimiser step against torch (clip + Adam)."""
import pytest
import torch

from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P

pytestmark = pytest.mark.gpu
KW = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
          n_repeats=1, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0)


def oracle_grads(model, kw, fusion, mix, ref, Tr, w_est, w_log):
    sd = {k: v.detach().clone().double() for k, v in model.state_dict().items()}
    leaves = {}
    for n, p in model.named_parameters():
        if p.requires_grad:
            sd[n] = sd[n].requires_grad_(True)
            leaves[n] = sd[n]
    cfg = O.Config(**{k: kw[k] for k in ('input_size', 'feature_size', 'hidden_size', 'chunk_length', 'kernel_size',
                                         'hop_length', 'n_repeats', 'bidirectional', 'norm_type', 'activation_type')},
                   fusion_type=fusion)
    est, logits = O.spe_forward(mix.double(), ref.double(), torch.tensor(float(Tr)), sd, cfg, training=True, new_stats={},
                                fast=False)
    loss = (est * w_est.double()).sum() + (logits * w_log.double()).sum()
    loss.backward()
    return est.detach(), logits.detach(), {n: v.grad for n, v in leaves.items()}


@pytest.mark.parametrize('fusion,extra', [('film', {}), ('cat', {}), ('add', {}), ('mul', dict(norm_type='gLN', bidirectional=False))])
def test_backward_matches_oracle_autograd(fusion, extra):
    kw = dict(KW, **extra)
    torch.manual_seed(11)
    model = P.DPRNNSpeTasNet(**kw, fusion_type=fusion).train()
    g = torch.Generator().manual_seed(12)
    B, T, Tr = 2, 1501, 1300
    mix, ref = 0.05 * torch.randn(B, T, generator=g), 0.05 * torch.randn(B, Tr, generator=g)
    w_est, w_log = torch.randn(B, T, generator=g), torch.randn(B, 251, generator=g)
    est_o, log_o, grads_o = oracle_grads(model, kw, fusion, mix, ref, Tr, w_est, w_log)
    model = model.cuda()
    est, logits = model(mix.cuda(), ref.cuda(), torch.tensor(float(Tr)))
    assert est.requires_grad and logits.requires_grad
    assert O.peak_rel_err(est.detach().cpu(), est_o.float()) < 2e-5
    loss = (est * w_est.cuda()).sum() + (logits * w_log.cuda()).sum()
    loss.backward()
    worst = ('', 0.0)
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        assert p.grad is not None, n
        want = grads_o[n].float()
        denom = max(float(want.abs().max()), 1e-6 * max(float(v.abs().max()) for v in grads_o.values()))
        err = float((p.grad.cpu() - want).abs().max()) / denom
        if err > worst[1]:
            worst = (n, err)
        assert err < 2e-3, (n, err)
    print('worst gradient error', worst)
