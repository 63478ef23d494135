"""How far torch's own fp32 autograd (through the CPU oracle) is from the fp64 gradient, next to the hand-written CUDA backward:
the noise floor the gradient tests of DPRNN-Spe-IRA are held to (tests/test_gpu_train.py).  GPU box: python tools/grad_noise_floor.py"""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P
from test_gpu_train import KW, _oracle_leaves
for cls, fwd in ((P.DPRNNSpeTasNet, O.spe_forward), (P.DPRNNSpeIRATasNet, O.ira_forward)):
    kw = dict(KW)
    torch.manual_seed(21)
    model = cls(**kw, fusion_type='cat').train()
    g = torch.Generator().manual_seed(22)
    B, T, Tr = 2, 1501, 1300
    mix, ref = 0.05 * torch.randn(B, T, generator=g), 0.05 * torch.randn(B, Tr, generator=g)
    w_est, w_log = torch.randn(B, T, generator=g), torch.randn(B, 251, generator=g)
    sd, leaves = _oracle_leaves(model)
    cfg = O.Config(n_repeats=1, fusion_type='cat')
    est_o, log_o = fwd(mix.double(), ref.double(), torch.tensor(float(Tr)), sd, cfg, training=True, new_stats={}, fast=False)
    ((est_o * w_est.double()).sum() + (log_o * w_log.double()).sum()).backward()
    # fp32 autograd through the same oracle: the noise floor of ANY fp32 evaluation of this gradient
    sd32 = {k: v.detach().float() for k, v in sd.items()}
    l32 = {}
    for n in leaves:
        sd32[n] = sd32[n].requires_grad_(True); l32[n] = sd32[n]
    e32, g32 = fwd(mix, ref, torch.tensor(float(Tr)), sd32, cfg, training=True, new_stats={}, fast=False)
    ((e32 * w_est).sum() + (g32 * w_log).sum()).backward()
    model = model.cuda()
    est, logits = model(mix.cuda(), ref.cuda(), torch.tensor(float(Tr)))
    ((est * w_est.cuda()).sum() + (logits * w_log.cuda()).sum()).backward()
    print(cls.__name__)
    rows = []
    for n, p in model.named_parameters():
        if not p.requires_grad: continue
        want = leaves[n].grad
        den = float(want.abs().max())
        ours = float((p.grad.cpu().double() - want).abs().max()) / den
        cpu32 = float((l32[n].grad.double() - want).abs().max()) / den
        rows.append((ours, cpu32, n))
    for ours, cpu32, n in sorted(rows, reverse=True)[:8]:
        print(f'   ours {ours:.2e}   torch-fp32-autograd {cpu32:.2e}   {n}')
