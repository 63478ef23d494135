"""Per-kernel parity: each C-ABI entry point against the CPU oracle on seeded inputs (B200 only)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P
from tss_with_dprnn_b200.engine import Engine, EPI_GATED, EPI_NONE, EPI_RELU, EPI_SIGMOID

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def L_():
    return P.lib()


def stream():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(*shape, generator=g)


def cl(x):      # [B,C,L] -> channels-last [B,L,C] on the GPU
    return x.permute(0, 2, 1).contiguous().to(DEV)


def test_build_info_loaded():
    assert 'sm_100a' in L_().build_info()
    assert torch.cuda.get_device_capability()[0] == 10


@pytest.mark.parametrize('B,T', [(2, 4000), (1, 24000), (3, 2)])
def test_encoder(B, T):
    x, w = rnd(B, T, seed=1, scale=0.05), rnd(64, 1, 2, seed=2)
    want = O.encoder(x, w)
    enc = torch.empty(B, T - 1, 64, device=DEV)
    L_().call('dprnn_encoder_fwd', x.to(DEV), w.reshape(64, 2).contiguous().to(DEV), enc, B, T, 64, 2, 1, stream())
    assert O.peak_rel_err(enc.cpu(), want.permute(0, 2, 1)) < 1e-6


@pytest.mark.parametrize('L', [23999, 3999, 30159, 249, 1])
def test_unfold_bit_exact_and_golden_map(L):
    z = np.load(os.path.join(GOLDEN, 'index_maps.npz'))
    K, Pp, F, B = 250, 125, 4, 2
    S = L_().query('dprnn_num_chunks', L, K, Pp)
    assert S == z[f'unfold_{L}'].shape[1]
    # frame t of utterance b is stored as (t+1) + b*1e6 in every channel -> recover the integer map
    y = (torch.arange(1, L + 1, dtype=torch.float32).view(1, L, 1) + 1e6 * torch.arange(B).view(B, 1, 1)).expand(B, L, F)
    x = torch.empty(B, S, K, F, device=DEV)
    L_().call('dprnn_unfold', y.contiguous().to(DEV), x, B, L, K, Pp, F, stream())
    x = x.cpu()
    for b in range(B):
        got = x[b, :, :, 0].t().long()                      # [K,S]
        got = torch.where(got == 0, torch.full_like(got, -1), got - 1 - b * 1000000)
        assert np.array_equal(got.numpy(), z[f'unfold_{L}']), (L, b)
    assert torch.equal(x[..., 0], x[..., 3])


@pytest.mark.parametrize('L', [3999, 24000, 1, 249])
def test_unfold_fold_vs_oracle_bit_exact(L):
    K, Pp, F, B = 250, 125, 128, 2
    y = rnd(B, F, L, seed=L)
    want = O.segmentation(y, K, Pp)                          # [B,F,K,S]
    S = want.shape[-1]
    x = torch.empty(B, S, K, F, device=DEV)
    L_().call('dprnn_unfold', cl(y), x, B, L, K, Pp, F, stream())
    assert torch.equal(x.cpu(), want.permute(0, 3, 2, 1))
    # fold of arbitrary chunk data (not just unfolded data) + PReLU
    c = rnd(B, F, K, S, seed=L + 1)
    a = torch.tensor([0.25])
    want_f = O.overlap_add(torch.where(c >= 0, c, a * c), L, K, Pp)
    out = torch.empty(B, L, F, device=DEV)
    L_().call('dprnn_fold_prelu', c.permute(0, 3, 2, 1).contiguous().to(DEV), out, B, L, K, Pp, F, a.to(DEV), stream())
    assert torch.equal(out.cpu(), want_f.permute(0, 2, 1))   # two-term sums: order-independent, bit-exact
    # round trip: fold(unfold(y)) == 2*y exactly
    L_().call('dprnn_fold_prelu', x, out, B, L, K, Pp, F, None, stream())
    assert torch.equal(out.cpu(), 2 * y.permute(0, 2, 1))


@pytest.mark.parametrize('eps', [1e-5, 1e-8])
def test_stats_norm_residual(eps):
    B, R, C = 3, 1237, 128
    y, x = rnd(B, C, R, seed=3) * 3 + 0.7, rnd(B, C, R, seed=4)
    g, b = rnd(C, seed=5), rnd(C, seed=6)
    want = x + O.chan_norm(y, g, b, eps)
    eng = Engine.__new__(Engine)
    yd, xd = cl(y), cl(x)
    mr = Engine.utt_stats(eng, yd, B, R * C, eps)
    mean = y.reshape(B, -1).mean(1)
    rstd = 1 / torch.sqrt(y.reshape(B, -1).var(1, unbiased=False) + eps)
    assert torch.allclose(mr[:, 0].cpu(), mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(mr[:, 1].cpu(), rstd, rtol=1e-5)
    L_().call('dprnn_norm_residual', yd, xd, mr, g.to(DEV), b.to(DEV), B, R, C, None, stream())
    assert O.peak_rel_err(xd.cpu(), want.permute(0, 2, 1)) < 2e-6


@pytest.mark.parametrize('M,N,K,epi', [(1000, 128, 64, EPI_NONE), (777, 1024, 128, EPI_NONE), (513, 128, 256, EPI_NONE),
                                        (64, 64, 128, EPI_SIGMOID), (130, 64, 128, EPI_RELU), (300, 256, 128, EPI_GATED),
                                        (5, 128, 128, EPI_NONE), (2000, 256, 128, EPI_NONE), (999, 128, 192, EPI_NONE)])
def test_gemm_f32(M, N, K, epi):
    A, Wm, bias = rnd(M, K, seed=M), rnd(N, K, seed=N + 1) / K ** 0.5, rnd(N, seed=7)
    ref = (A.double() @ Wm.double().t() + bias.double())
    if epi == EPI_SIGMOID:
        ref = torch.sigmoid(ref)
    elif epi == EPI_RELU:
        ref = torch.relu(ref)
    if epi == EPI_GATED:
        H = N // 2
        ref = torch.tanh(ref[:, :H]) * torch.sigmoid(ref[:, H:])
        cols = []
        bcols = []
        for t0 in range(0, H, 64):
            cols += [Wm[t0:t0 + 64].t(), Wm[H + t0:H + t0 + 64].t()]
            bcols += [bias[t0:t0 + 64], bias[H + t0:H + t0 + 64]]
        Wt, bb = torch.cat(cols, 1).contiguous(), torch.cat(bcols)
    else:
        Wt, bb = Wm.t().contiguous(), bias
    eng = Engine.__new__(Engine)
    out = Engine.gemm(eng, A.to(DEV), Wt.to(DEV), M, N, K, bias=bb.to(DEV), epi=epi)
    assert O.peak_rel_err(out.cpu(), ref.float()) < 2e-6


def test_gemm_prologue_and_per_utt_bias():
    B, R, K, N = 3, 517, 64, 128
    A = rnd(B * R, K, seed=1)
    Wm = rnd(N, K, seed=2) / 8
    s1, s0, add = rnd(B, K, seed=3), rnd(B, K, seed=4), rnd(B, K, seed=5)
    rs, bias = rnd(B * R, seed=6), rnd(B, N, seed=7)
    a3 = A.view(B, R, K).double()
    pro = (a3 * s1[:, None].double() + s0[:, None].double()) * rs.view(B, R, 1).double() + add[:, None].double()
    ref = pro @ Wm.double().t() + 2.0 * bias[:, None].double()
    eng = Engine.__new__(Engine)
    out = Engine.gemm(eng, A.to(DEV), Wm.t().contiguous().to(DEV), B * R, N, K, bias=bias.to(DEV), bias_per_utt=True,
                      bias_scale=2.0, rows_per_utt=R, p_scale=s1.to(DEV), p_shift=s0.to(DEV), p_add=add.to(DEV),
                      rowscale=rs.to(DEV))
    assert O.peak_rel_err(out.cpu(), ref.view(B * R, N).float()) < 2e-6


@pytest.mark.parametrize('geom', ['intra', 'inter'])
@pytest.mark.parametrize('ndir', [2, 1])
def test_lstm_recurrence_f32(geom, ndir):
    H = 128
    B, S, K = 2, 7, 19                      # tiny chunk grid: rows = B*S*K
    rows = B * S * K
    torch.manual_seed(11)
    lstm = torch.nn.LSTM(H, H, batch_first=True, bidirectional=(ndir == 2))
    sd = {'r.' + k: v.detach() for k, v in lstm.state_dict().items()}
    x = rnd(B, S, K, H, seed=12)
    if geom == 'intra':
        seqs = x.reshape(B * S, K, H)
        geo = (B * S, K, 1, K, 0, 1)
    else:
        seqs = x.permute(0, 2, 1, 3).reshape(B * K, S, H)
        geo = (B * K, S, K, S * K, 1, K)
    want = O.lstm(seqs, sd, 'r', ndir == 2, fast=False)                       # [nseq, T, nd*H]
    want = want.reshape(B, S, K, ndir * H) if geom == 'intra' else want.reshape(B, K, S, ndir * H).permute(0, 2, 1, 3)
    sfx = ['', '_reverse'][:ndir]
    wih = torch.cat([sd['r.weight_ih_l0' + s] for s in sfx], 0)
    bias = torch.cat([sd['r.bias_ih_l0' + s] + sd['r.bias_hh_l0' + s] for s in sfx], 0)
    whh_t = torch.stack([sd['r.weight_hh_l0' + s].t() for s in sfx], 0).contiguous().to(DEV)
    eng = Engine.__new__(Engine)
    gx = Engine.gemm(eng, x.reshape(rows, H).to(DEV), wih.t().contiguous().to(DEV), rows, ndir * 4 * H, H, bias=bias.to(DEV))
    hout = torch.empty(rows, ndir * H, device=DEV)
    L_().call('dprnn_lstm_recurrence_f32', gx, whh_t, hout, *geo, H, ndir, stream())
    assert O.peak_rel_err(hout.cpu().view(B, S, K, ndir * H), want) < 5e-6


def test_lstm_recurrence_many_sequences():
    """More sequences than one CTA tile, not a multiple of 32; checks tile edges."""
    H, nseq, T = 128, 77, 33
    torch.manual_seed(13)
    lstm = torch.nn.LSTM(H, H, batch_first=True, bidirectional=True)
    sd = {'r.' + k: v.detach() for k, v in lstm.state_dict().items()}
    x = rnd(nseq, T, H, seed=14)
    want = O.lstm(x, sd, 'r', True, fast=True)
    sfx = ['', '_reverse']
    wih = torch.cat([sd['r.weight_ih_l0' + s] for s in sfx], 0)
    bias = torch.cat([sd['r.bias_ih_l0' + s] + sd['r.bias_hh_l0' + s] for s in sfx], 0)
    whh_t = torch.stack([sd['r.weight_hh_l0' + s].t() for s in sfx], 0).contiguous().to(DEV)
    eng = Engine.__new__(Engine)
    gx = Engine.gemm(eng, x.reshape(-1, H).to(DEV), wih.t().contiguous().to(DEV), nseq * T, 8 * H, H, bias=bias.to(DEV))
    hout = torch.empty(nseq * T, 2 * H, device=DEV)
    L_().call('dprnn_lstm_recurrence_f32', gx, whh_t, hout, nseq, T, 1, T, 0, 1, H, 2, stream())
    assert O.peak_rel_err(hout.cpu().view(nseq, T, 2 * H), want) < 5e-6


@pytest.mark.parametrize('B,L', [(2, 3999), (1, 1)])
def test_mask_decode(B, L):
    N = 64
    m, e, w = torch.rand(B, N, L, generator=torch.Generator().manual_seed(1)), rnd(B, N, L, seed=2).abs(), rnd(N, 1, 2, seed=3)
    want = O.decoder(m * e, w)
    out = torch.empty(B, L + 1, device=DEV)
    L_().call('dprnn_mask_decode', cl(m), L * N, cl(e), w.reshape(N, 2).contiguous().to(DEV), out, L + 1, B, L, N, 2, 1, stream())
    assert O.peak_rel_err(out.cpu(), want) < 2e-6
    d0 = torch.empty(B, L, N, device=DEV)
    L_().call('dprnn_mask_apply', cl(m), cl(e), d0, B * L * N, stream())
    assert torch.equal(d0.cpu(), (m * e).permute(0, 2, 1))


@pytest.mark.parametrize('L', [3999, 4000, 5, 4])
def test_attention_rowscale(L):
    B, N, k = 2, 64, 2
    x = rnd(B, N, L, seed=L).abs()
    g, b = rnd(N, seed=1), rnd(N, seed=2)
    v = rnd(B, N, seed=3)
    wavg, bavg = torch.full((N, 1, k), 0.5), torch.zeros(N)
    xn = O.chan_norm(x, g, b, 1e-5)
    La = (L - k) // k + 1
    avg = (xn[..., :La * k].reshape(B, N, La, k) * wavg.view(1, N, 1, k)).sum(-1) + bavg.view(1, N, 1)
    sm = torch.softmax((avg * v.unsqueeze(-1)).sum(1), -1)
    src = torch.from_numpy(O.nearest_upsample_index(La, L))
    want = 1 + sm[:, src]
    eng = Engine.__new__(Engine)
    xd = cl(x)
    mr = Engine.utt_stats(eng, xd, B, L * N, 1e-5)
    n1, n0 = torch.empty(B, N, device=DEV), torch.empty(B, N, device=DEV)
    L_().call('dprnn_norm_affine', mr, g.to(DEV), b.to(DEV), None, n1, n0, B, N, stream())
    scores, rowscale = torch.empty(B, La, device=DEV), torch.empty(B, L, device=DEV)
    L_().call('dprnn_att_rowscale', xd, n1, n0, wavg.reshape(N, k).contiguous().to(DEV), bavg.to(DEV), v.to(DEV), scores,
              rowscale, B, L, N, k, stream())
    assert torch.allclose(rowscale.cpu(), want, rtol=2e-5, atol=1e-7)
    # the index map itself, bit-exact against the ATen fixture: feed a softmax that encodes the position
    z = np.load(os.path.join(GOLDEN, 'index_maps.npz'))
    key = f'nearest_{L}'
    if key in z.files:
        assert np.array_equal(src.numpy(), z[key])


def test_small_linear_and_time_sum():
    B, E, N = 3, 128, 251
    lin = torch.nn.Linear(E, N)
    x = rnd(B, E, seed=1)
    eng = Engine.__new__(Engine)
    lin_d = torch.nn.Linear(E, N).to(DEV)
    lin_d.load_state_dict(lin.state_dict())
    out = Engine.small_linear(eng, x.to(DEV), lin_d, B)
    assert O.peak_rel_err(out.cpu(), lin(x).detach()) < 2e-6
    z = rnd(B, 888, 128, seed=2)
    div = torch.tensor([888., 100., 7.])
    emb = torch.empty(B, 128, device=DEV)
    L_().call('dprnn_time_sum', z.to(DEV), emb, B, 888, 128, div.to(DEV), stream())
    assert O.peak_rel_err(emb.cpu(), z.sum(1) / div[:, None]) < 2e-6
