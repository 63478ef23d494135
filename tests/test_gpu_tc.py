"""Tensor-core (tcgen05 + TMA) kernels against fp64 references computed from the same bf16-rounded inputs."""
import pytest
import torch

from oracle import dprnn_oracle as O
import tss_with_dprnn_b200 as P

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def stream():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize('M,N,K', [(128, 128, 256), (1000, 128, 256), (48500, 128, 256), (300, 128, 128),
                                   (257, 64, 128), (129, 256, 128)])
def test_linear_bf16(M, N, K):
    A = rnd(M, K, seed=M).bfloat16()
    W = (rnd(N, K, seed=N) / K ** 0.5).bfloat16()
    bias = rnd(N, seed=3)
    ref = (A.double() @ W.double().t() + bias.double()).float()
    out = torch.full((M, N), float('nan'), device=DEV)
    P.lib().call('dprnn_linear_bf16', A.to(DEV), W.to(DEV), bias.to(DEV), out, N, M, N, K, stream())
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert O.peak_rel_err(out.cpu(), ref) < 1e-5      # exact products, fp32 accumulation


def lstm_bf16_reference(x, rnn, reverse_flags, exact_h_rounding=True, dtype=torch.bfloat16):
    """fp32 restatement of what dprnn_lstm_layer_bf16 computes: bf16 (or fp16) x / weights / recurrent h, fp32
    accumulation and state, exact activations. x [nseq, T, 128] (already representable in the 16-bit format)."""
    outs = []
    for d, rev in enumerate(reverse_flags):
        sf = '_reverse' if rev else ''
        wih = getattr(rnn, 'weight_ih_l0' + sf).detach().to(dtype).float()
        whh = getattr(rnn, 'weight_hh_l0' + sf).detach().to(dtype).float()
        b = (getattr(rnn, 'bias_ih_l0' + sf) + getattr(rnn, 'bias_hh_l0' + sf)).detach()
        N, T, H = x.shape[0], x.shape[1], 128
        h = torch.zeros(N, H); c = torch.zeros(N, H)
        out = torch.empty(N, T, H)
        for t in (range(T - 1, -1, -1) if rev else range(T)):
            g = x[:, t] @ wih.t() + h.to(dtype).float() @ whh.t() + b
            i, f, gg, o = g.split(H, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            out[:, t] = h
        outs.append(out)
    return torch.cat(outs, -1)


@pytest.mark.parametrize('inter', [0, 1])
@pytest.mark.parametrize('ndir', [2, 1])
@pytest.mark.parametrize('fast', [0, 1])
def test_lstm_layer_bf16(inter, ndir, fast):
    from tss_with_dprnn_b200.engine import Engine
    H = 128
    # intra: 300 sequences (more than one 256-sequence pair tile, ragged tail), T = 37
    # inter: K = 250 sequences per utterance (one pair tile per utterance, 6 rows of zero-filled padding), T = 21
    B, S, K = (3, 100, 37) if not inter else (3, 21, 250)
    rows = B * S * K
    torch.manual_seed(21 + inter)
    rnn = torch.nn.LSTM(H, H, batch_first=True, bidirectional=(ndir == 2))
    x = rnd(B, S, K, H, seed=22).bfloat16()
    xf = x.float()
    seqs = xf.reshape(B * S, K, H) if not inter else xf.permute(0, 2, 1, 3).reshape(B * K, S, H)
    want = lstm_bf16_reference(seqs, rnn, [False, True][:ndir])
    want = want.reshape(B, S, K, ndir * H) if not inter else want.reshape(B, K, S, ndir * H).permute(0, 2, 1, 3)
    wp, bp = Engine._pack_lstm_tc(rnn, ['', '_reverse'][:ndir])
    hout = torch.full((rows, ndir * H), float('nan'), device=DEV, dtype=torch.bfloat16)
    P.lib().call('dprnn_lstm_layer_bf16', x.reshape(rows, H).to(DEV), wp.to(DEV), bp.to(DEV), hout, B, S, K, inter, H, ndir,
                 fast, stream())
    torch.cuda.synchronize()
    got = hout.float().cpu().view(B, S, K, ndir * H)
    assert torch.isfinite(got).all()
    err = O.peak_rel_err(got, want)
    # bf16 output rounding (2^-9) + rounding flips of the recurrent h; tanh.approx adds ~2^-11
    assert err < (1.5e-2 if fast else 1e-2), err
    assert (got - want).abs().mean() < 2e-3


@pytest.mark.parametrize('M,N,K,bf16,epi', [(1000, 128, 256, True, 0), (48500 * 2, 128, 256, True, 0), (300, 128, 128, True, 0),
                                            (777, 128, 128, False, 0), (1000, 256, 128, False, 0), (513, 256, 256, False, 0),
                                            (129, 128, 256, False, 0), (640, 64, 128, False, 2), (640, 64, 128, False, 1),
                                            (900, 256, 128, False, 3), (5, 128, 128, False, 0)])
def test_gemm_tc(M, N, K, bf16, epi):
    from tss_with_dprnn_b200.engine import Engine
    A = rnd(M, K, seed=M)
    W = rnd(N, K, seed=N + 1) / K ** 0.5
    bias = rnd(N, seed=3)
    if bf16:
        A, W = A.bfloat16(), W.bfloat16()
        Ar, Wr = A.double(), W.double()
    else:   # TF32: operands truncated to 10 mantissa bits
        Ar = (A.view(torch.int32) & ~0x1FFF).view(torch.float32).double()
        Wr = (W.view(torch.int32) & ~0x1FFF).view(torch.float32).double()
    ref = Ar @ Wr.t() + bias.double()
    if epi == 1:
        ref = torch.relu(ref)
    elif epi == 2:
        ref = torch.sigmoid(ref)
    elif epi == 3:
        ref = torch.tanh(ref[:, :N // 2]) * torch.sigmoid(ref[:, N // 2:])
    eng = Engine.__new__(Engine)
    out = Engine.gemm_tc(eng, A.to(DEV), W.to(DEV), M, N, K, bias=bias.to(DEV), epi=epi)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    tol = 1e-5 if epi in (0, 1) else 2e-3          # tanh.approx-based activations in the epilogue
    assert O.peak_rel_err(out.cpu(), ref.float()) < tol


def test_gemm_tc_fused_stats():
    from tss_with_dprnn_b200.engine import Engine
    B, R, K, N = 3, 1337, 256, 128           # tiles straddle utterance boundaries (1337 % 128 != 0)
    A = (rnd(B * R, K, seed=1) + 0.3).bfloat16()
    W = (rnd(N, K, seed=2) / 16).bfloat16()
    bias = rnd(N, seed=3)
    ref = (A.double() @ W.double().t() + bias.double()).view(B, -1)
    eng = Engine.__new__(Engine)
    out, mr = Engine.gemm_tc(eng, A.to(DEV), W.to(DEV), B * R, N, K, bias=bias.to(DEV), stats=(R, 1e-5))
    mean, rstd = ref.mean(1), 1 / torch.sqrt(ref.var(1, unbiased=False) + 1e-5)
    assert torch.allclose(mr[:, 0].cpu().double(), mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(mr[:, 1].cpu().double(), rstd, rtol=1e-4)
    out2, mr2 = Engine.gemm_tc(eng, A.to(DEV), W.to(DEV), B * R, N, K, bias=bias.to(DEV), stats=(R, 1e-5))
    assert torch.equal(mr, mr2)            # deterministic reduction


@pytest.mark.parametrize('B,R,K', [(3, 1337, 256), (1, 48500, 256), (2, 700, 128), (64, 300, 256)])
def test_linear_persistent(B, R, K):
    M, N = B * R, 128
    A = (rnd(M, K, seed=M) + 0.2).bfloat16()
    W = (rnd(N, K, seed=5) / K ** 0.5).bfloat16()
    bias = rnd(N, seed=3)
    ref = A.double() @ W.double().t() + bias.double()
    out = torch.full((M, N), float('nan'), device=DEV)
    part = torch.empty(P.lib().query('dprnn_gemm_tc_stats_bytes', M), device=DEV, dtype=torch.uint8)
    mr = torch.empty(B, 2, device=DEV)
    P.lib().call('dprnn_linear_bf16_stats', A.to(DEV), W.to(DEV), bias.to(DEV), out, M, K, part, R, 1e-5, mr, stream())
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert O.peak_rel_err(out.cpu(), ref.float()) < 1e-5
    rb = ref.view(B, -1)
    assert torch.allclose(mr[:, 0].cpu().double(), rb.mean(1), rtol=1e-4, atol=1e-5)
    assert torch.allclose(mr[:, 1].cpu().double(), 1 / torch.sqrt(rb.var(1, unbiased=False) + 1e-5), rtol=1e-4)


@pytest.mark.parametrize('lens,K', [([1337, 1337, 1337], 256), ([48500], 256), ([700, 700], 128), ([300] * 64, 256),
                                    ([750, 2250, 1000, 48500, 1250], 256), ([24250] * 9, 256)])
def test_linear_norm_residual_fused(lens, K):
    """Linear + GroupNorm(1,128) statistics + apply + residual in one persistent kernel (linear_norm.cu) against an fp64
    restatement of dprnn.py:86-92; utterances of unequal length that do not align with the 128-row tiles."""
    M, N, n = sum(lens), 128, len(lens)
    h = (rnd(M, K, seed=M % 1000) + 0.2).bfloat16()
    W = (rnd(N, K, seed=5) / K ** 0.5).bfloat16()
    bias, gamma, beta = rnd(N, seed=3), 1 + 0.1 * rnd(N, seed=4), 0.1 * rnd(N, seed=6)
    x0 = rnd(M, N, seed=7)
    y = h.double() @ W.double().t() + bias.double()
    want = x0.double().clone()
    off = [0]
    for ln in lens:
        off.append(off[-1] + ln)
    for u in range(n):
        yu = y[off[u]:off[u + 1]]
        mean, var = yu.mean(), yu.var(unbiased=False)
        want[off[u]:off[u + 1]] += (yu - mean) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()
    L = P.lib()
    ws = torch.empty(L.query('dprnn_linear_norm_workspace_bytes', M, n), device=DEV, dtype=torch.uint8)
    outs = []
    for _ in range(2):
        x = x0.clone().to(DEV)
        xb = torch.full((M, N), float('nan'), device=DEV, dtype=torch.bfloat16)
        L.call('dprnn_linear_norm_residual_bf16', h.to(DEV), W.to(DEV), bias.to(DEV), x, xb, gamma.to(DEV), beta.to(DEV),
               1e-5, torch.tensor(off, dtype=torch.int64, device=DEV), n, max(lens), M, K, ws, stream())
        torch.cuda.synchronize()
        outs.append((x.cpu(), xb.cpu()))
    x, xb = outs[0]
    assert torch.isfinite(x).all()
    assert O.peak_rel_err(x, want.float()) < 2e-5
    assert torch.equal(xb, x.bfloat16())                      # the shadow copy is the rounded new x
    assert torch.equal(outs[1][0], x)                         # deterministic statistics


@pytest.mark.parametrize('B,R,K', [(3, 1337, 256), (1, 48500, 256), (2, 700, 128)])
def test_linear_bf16out_then_norm_residual(B, R, K):
    """Linear with bf16 output + fp32-accurate statistics, then norm + residual reading the bf16 y (dprnn.py:86-92)."""
    M, N = B * R, 128
    A = (rnd(M, K, seed=M) + 0.2).bfloat16()
    W = (rnd(N, K, seed=5) / K ** 0.5).bfloat16()
    bias, gamma, beta = rnd(N, seed=3), 1 + 0.1 * rnd(N, seed=4), 0.1 * rnd(N, seed=6)
    x0 = rnd(M, N, seed=7)
    ref = A.double() @ W.double().t() + bias.double()
    L = P.lib()
    y = torch.full((M, N), float('nan'), device=DEV, dtype=torch.bfloat16)
    part = torch.empty(L.query('dprnn_gemm_tc_stats_bytes', M), device=DEV, dtype=torch.uint8)
    mr = torch.empty(B, 2, device=DEV)
    L.call('dprnn_linear_bf16out_stats', A.to(DEV), W.to(DEV), bias.to(DEV), y, M, K, part, R, 1e-5, mr, stream())
    torch.cuda.synchronize()
    assert torch.equal(y.cpu(), ref.float().bfloat16()) or O.peak_rel_err(y.float().cpu(), ref.float()) < 4e-3
    rb = ref.view(B, -1)
    mean, rstd = rb.mean(1), 1 / torch.sqrt(rb.var(1, unbiased=False) + 1e-5)
    assert torch.allclose(mr[:, 0].cpu().double(), mean, rtol=1e-4, atol=1e-5)      # statistics are NOT from the rounded y
    assert torch.allclose(mr[:, 1].cpu().double(), rstd, rtol=1e-4)
    x = x0.clone().to(DEV)
    xb = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
    L.call('dprnn_norm_residual_ybf16', y, x, mr, gamma.to(DEV), beta.to(DEV), B, R, N, xb, stream())
    torch.cuda.synchronize()
    yv = y.float().cpu().double().view(B, R, N)
    want = x0.double().view(B, R, N) + (yv - mean.view(B, 1, 1)) * rstd.view(B, 1, 1) * gamma.double() + beta.double()
    assert O.peak_rel_err(x.cpu(), want.view(M, N).float()) < 2e-5
    assert torch.equal(xb.cpu(), x.cpu().bfloat16())


@pytest.mark.parametrize('inter', [0, 1])
@pytest.mark.parametrize('nslices', [0, 1, 2, 3, 7])
@pytest.mark.parametrize('fmt', ['bf16', 'fp16'])
def test_lstm_sliced_persistent_is_bit_identical(inter, nslices, fmt):
    """The persistent, time-sliced half-job LSTM kernel (cell state through a global scratch, h re-read from the stored
    rows, weights re-loaded when the drawn item changes direction) returns bit for bit what the one-job-per-pair half-job
    kernel returns; nslices = 0 lets the launcher choose; 3 resident pairs force many items per pair."""
    from tss_with_dprnn_b200.engine import Engine
    L = P.lib()
    B, S, K, H, nd = 5, 30, 250, 128, 2                      # intra: 150 sequences; inter: 5 tiles x 2 dirs, T = 30
    dtype = torch.float16 if fmt == 'fp16' else torch.bfloat16
    flags = 1 | (2 if fmt == 'fp16' else 0)
    torch.manual_seed(3 + inter)
    rnn = torch.nn.LSTM(H, H, 1, batch_first=True, bidirectional=True).cuda()
    rows = B * S * K
    xb = (0.5 * torch.randn(rows, H, generator=torch.Generator().manual_seed(4))).cuda().to(dtype)
    wp2, bp = Engine._pack_lstm_tc(rnn, ['', '_reverse'], half_jobs=True, dtype=dtype)
    st = torch.cuda.current_stream().cuda_stream
    want = torch.empty(rows, nd * H, device='cuda', dtype=dtype)
    L.call('dprnn_lstm_layer_bf16_pp', xb, wp2, bp, want, B, S, K, inter, H, nd, flags, st)
    got = torch.full((rows + 1, nd * H), 7.0, device='cuda', dtype=dtype)
    ws = torch.empty(L.query('dprnn_lstm_sliced_workspace_bytes', B, S, K, inter, nd), device='cuda', dtype=torch.uint8)
    for _ in range(2):                                       # the workspace is re-armed by every call
        L.call('dprnn_lstm_layer_bf16_sliced', xb, wp2, bp, got, B, S, K, inter, H, nd, flags, nslices,
               3 if nslices in (0, 3) else 0, ws, st)
        torch.cuda.synchronize()
        assert torch.equal(got[:rows], want)
        assert float((got[rows:].float() - 7.0).abs().max()) == 0.0


def test_lstm_sliced_auto_picks_the_packing_slices():
    """dprnn_lstm_sliced_auto minimises ceil(k J / pairs) * ceil(T / k): the headline layers (B = 64, 3 s) on 74 pairs."""
    L = P.lib()
    assert L.query('dprnn_lstm_sliced_auto', 64, 194, 250, 0, 2, 74, 8) == 3      # intra: J = 98, 294 items = 3.97 rounds
    k_inter = L.query('dprnn_lstm_sliced_auto', 64, 194, 250, 1, 2, 74, 8)        # inter: J = 128
    import math
    cost = lambda k, J, T: math.ceil(k * J / 74) * math.ceil(T / k)
    assert cost(k_inter, 128, 194) == min(cost(k, 128, 194) for k in range(1, 9))
    assert L.query('dprnn_lstm_sliced_auto', 1, 194, 250, 0, 2, 74, 8) == 1       # one wave already: no slicing


@pytest.mark.parametrize('inter', [0, 1])
def test_lstm_half_job_pingpong_matches(inter):
    """The half-job ping-pong kernel (M = 128 MMAs, 2x2 accumulator layout, re-packed weight rows) against the one-job
    kernel on the same inputs."""
    from tss_with_dprnn_b200.engine import Engine
    L = P.lib()
    B, S, K, H, nd = 5, 30, 250, 128, 2
    torch.manual_seed(5 + inter)
    rnn = torch.nn.LSTM(H, H, 1, batch_first=True, bidirectional=True).cuda()
    rows = B * S * K
    xb = (0.5 * torch.randn(rows, H, generator=torch.Generator().manual_seed(6))).cuda().to(torch.bfloat16)
    wp, bp = Engine._pack_lstm_tc(rnn, ['', '_reverse'])
    wp2, bp2 = Engine._pack_lstm_tc(rnn, ['', '_reverse'], half_jobs=True)
    assert torch.equal(bp, bp2)
    st = torch.cuda.current_stream().cuda_stream
    want = torch.empty(rows, nd * H, device='cuda', dtype=torch.bfloat16)
    L.call('dprnn_lstm_layer_bf16', xb, wp, bp, want, B, S, K, inter, H, nd, 1, st)
    got = torch.full((rows + 1, nd * H), 7.0, device='cuda', dtype=torch.bfloat16)
    L.call('dprnn_lstm_layer_bf16_pp', xb, wp2, bp2, got, B, S, K, inter, H, nd, 1, st)
    torch.cuda.synchronize()
    err = float((got[:rows].float() - want.float()).abs().max())
    print('half-job kernel: max |diff| vs one-job kernel', err, 'bit-identical', torch.equal(got[:rows], want))
    assert err < 2e-2
    assert float((got[rows:].float() - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize('inter', [0, 1])
@pytest.mark.parametrize('ndir', [2, 1])
@pytest.mark.parametrize('fast', [0, 1])
@pytest.mark.parametrize('fmt', ['bf16', 'fp16'])
def test_lstm_layer_pp_direct(inter, ndir, fast, fmt):
    """The SHIPPED default LSTM kernel (dprnn_lstm_layer_bf16_pp, half-job ping-pong) directly against the fp32
    restatement of an nn.LSTM layer (dprnn.py:23-28) with the same operand roundings, in both 16-bit operand formats;
    in bf16 additionally bit for bit against the one-job kernel (DESIGN.md 4.1 states they are identical)."""
    from tss_with_dprnn_b200.engine import Engine
    H = 128
    dtype = torch.float16 if fmt == 'fp16' else torch.bfloat16
    B, S, K = (3, 100, 37) if not inter else (3, 21, 250)
    rows = B * S * K
    torch.manual_seed(21 + inter)
    rnn = torch.nn.LSTM(H, H, batch_first=True, bidirectional=(ndir == 2))
    x = rnd(B, S, K, H, seed=22).to(dtype)
    xf = x.float()
    seqs = xf.reshape(B * S, K, H) if not inter else xf.permute(0, 2, 1, 3).reshape(B * K, S, H)
    want = lstm_bf16_reference(seqs, rnn, [False, True][:ndir], dtype=dtype)
    want = want.reshape(B, S, K, ndir * H) if not inter else want.reshape(B, K, S, ndir * H).permute(0, 2, 1, 3)
    sfx = ['', '_reverse'][:ndir]
    wp2, bp = Engine._pack_lstm_tc(rnn, sfx, half_jobs=True, dtype=dtype)
    hout = torch.full((rows + 1, ndir * H), 7.0, device=DEV, dtype=dtype)
    flags = fast | (2 if fmt == 'fp16' else 0)
    xd = x.reshape(rows, H).to(DEV)
    P.lib().call('dprnn_lstm_layer_bf16_pp', xd, wp2.to(DEV), bp.to(DEV), hout, B, S, K, inter, H, ndir, flags, stream())
    torch.cuda.synchronize()
    got = hout[:rows].float().cpu().view(B, S, K, ndir * H)
    assert torch.isfinite(got).all()
    assert float((hout[rows:].float() - 7.0).abs().max()) == 0.0          # nothing written past the last row
    err = O.peak_rel_err(got, want)
    # output rounding (2^-9 bf16 / 2^-12 fp16) + rounding flips of the recurrent h; tanh.approx adds ~2^-11
    tol = {('bf16', 0): 1e-2, ('bf16', 1): 1.5e-2, ('fp16', 0): 1.5e-3, ('fp16', 1): 3e-3}[(fmt, fast)]
    assert err < tol, err
    assert (got - want).abs().mean() < (2e-3 if fmt == 'bf16' else 3e-4)
    if fmt == 'bf16':
        wp, _ = Engine._pack_lstm_tc(rnn, sfx)
        one = torch.empty((rows, ndir * H), device=DEV, dtype=dtype)
        P.lib().call('dprnn_lstm_layer_bf16', xd, wp.to(DEV), bp.to(DEV), one, B, S, K, inter, H, ndir, fast, stream())
        torch.cuda.synchronize()
        assert torch.equal(hout[:rows], one)


@pytest.mark.parametrize('B,R,K', [(3, 1337, 256), (2, 700, 128)])
@pytest.mark.parametrize('res16', [False, True])
def test_linear_fp16out_then_norm_residual(B, R, K, res16):
    """The half-block tail in the fp16 format: Linear (fp16 operands / output, fp32 statistics), then norm + residual
    with the fp32 master (yh16) or the 16-bit residual stream (h16res) - dprnn.py:86-92."""
    M, N = B * R, 128
    A = (rnd(M, K, seed=M) + 0.2).half()
    W = (rnd(N, K, seed=5) / K ** 0.5).half()
    bias, gamma, beta = rnd(N, seed=3), 1 + 0.1 * rnd(N, seed=4), 0.1 * rnd(N, seed=6)
    x0 = rnd(M, N, seed=7)
    ref = A.double() @ W.double().t() + bias.double()
    L = P.lib()
    y = torch.full((M, N), float('nan'), device=DEV, dtype=torch.float16)
    part = torch.empty(L.query('dprnn_gemm_tc_stats_bytes', M), device=DEV, dtype=torch.uint8)
    mr = torch.empty(B, 2, device=DEV)
    L.call('dprnn_linear_h16out_stats', A.to(DEV), W.to(DEV), bias.to(DEV), y, M, K, part, R, 1e-5, mr, 1, stream())
    torch.cuda.synchronize()
    assert O.peak_rel_err(y.float().cpu(), ref.float()) < 5e-4
    rb = ref.view(B, -1)
    mean, rstd = rb.mean(1), 1 / torch.sqrt(rb.var(1, unbiased=False) + 1e-5)
    assert torch.allclose(mr[:, 0].cpu().double(), mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(mr[:, 1].cpu().double(), rstd, rtol=1e-4)
    yv = y.float().cpu().double().view(B, R, N)
    nrm = (yv - mean.view(B, 1, 1)) * rstd.view(B, 1, 1) * gamma.double() + beta.double()
    if not res16:
        x = x0.clone().to(DEV)
        xb = torch.empty((M, N), device=DEV, dtype=torch.float16)
        L.call('dprnn_norm_residual_yh16', y, x, mr, gamma.to(DEV), beta.to(DEV), B, R, N, xb, 1, stream())
        torch.cuda.synchronize()
        want = x0.double().view(B, R, N) + nrm
        assert O.peak_rel_err(x.cpu(), want.view(M, N).float()) < 2e-5
        assert torch.equal(xb.cpu(), x.cpu().half())
    else:
        xb0 = x0.half()
        for last in (False, True):
            xb = xb0.clone().to(DEV)
            xf = torch.full((M, N), float('nan'), device=DEV)
            L.call('dprnn_norm_residual_h16res', y, xb, xf if last else None, mr, gamma.to(DEV), beta.to(DEV), B, R, N, 1,
                   stream())
            torch.cuda.synchronize()
            want = (xb0.double().view(B, R, N) + nrm).view(M, N).float()
            if last:                          # the fp32 result goes to x_f32_out, the 16-bit stream is left alone
                assert O.peak_rel_err(xf.cpu(), want) < 2e-5 and torch.equal(xb.cpu(), xb0)
            else:
                assert O.peak_rel_err(xb.float().cpu(), want) < 6e-4


@pytest.mark.parametrize('M,N,K,epi', [(777, 128, 128, 0), (1000, 256, 128, 0), (513, 256, 256, 0), (129, 128, 256, 0),
                                       (70000, 128, 64, 0), (640, 64, 128, 2), (900, 256, 128, 3), (5, 128, 128, 1),
                                       (3000, 128, 64, 4), (2048, 256, 256, 4)])
def test_gemm_f32x2_split(M, N, K, epi):
    """DPRNN_GEMM_F32X2: fp32 operands as bf16 pairs (in-kernel split of A, host-packed W, 3 MMAs): within 3e-5 of the
    fp64 product of the UNROUNDED operands, where the TF32 form (truncation) sits at ~1e-3.  N = K = 256 goes through
    the two-halves path of Engine.gemm_tc."""
    from tss_with_dprnn_b200.engine import Engine
    A = rnd(M, K, seed=M) + 0.25
    W = rnd(N, K, seed=N + 1) / K ** 0.5
    bias = rnd(N, seed=3)
    ref = A.double() @ W.double().t() + bias.double()
    eng = Engine.__new__(Engine)
    eng.conv_kind, eng.precision = 'f32x2', 'fp16'
    kw = dict(bias=bias.to(DEV), epi=epi)
    if epi == 1:
        ref = torch.relu(ref)
    elif epi == 2:
        ref = torch.sigmoid(ref)
    elif epi == 3:
        ref = torch.tanh(ref[:, :N // 2]) * torch.sigmoid(ref[:, N // 2:])
    elif epi == 4:
        sc, sh, pa = 1 + 0.1 * rnd(N, seed=5), 0.1 * rnd(N, seed=6), torch.tensor([0.25])
        ref = (A.double() @ W.double().t()) * sc.double() + sh.double()
        ref = torch.where(ref >= 0, ref, 0.25 * ref)
        kw = dict(post=(sc.to(DEV), sh.to(DEV), pa.to(DEV)))
    n0 = P.lib().launches
    out = Engine.gemm_tc(eng, A.to(DEV), W.to(DEV), M, N, K, **kw)
    torch.cuda.synchronize()
    assert P.lib().launches - n0 == (2 if N == 256 and K == 256 else 1)
    assert torch.isfinite(out).all()
    err = O.peak_rel_err(out.cpu(), ref.float())
    assert err < (3e-5 if epi in (0, 1, 4) else 1e-3), err           # tanh.approx in the sigmoid / gated epilogues
    eng.conv_kind = 'tf32'
    out_t = Engine.gemm_tc(eng, A.to(DEV), W.to(DEV), M, N, K, **kw)
    if epi == 0:
        assert O.peak_rel_err(out_t.cpu(), ref.float()) > 3 * err    # what the split buys


@pytest.mark.parametrize('inter', [0, 1])
@pytest.mark.parametrize('ndir', [2, 1])
@pytest.mark.parametrize('fmt', ['bf16', 'fp16'])
def test_lstm_fused_input_norm_is_bit_identical(inter, ndir, fmt):
    """dprnn_lstm_layer_bf16_pp_fused (the previous half-block's norm + residual applied while the LSTM kernel loads its
    input) == dprnn_norm_residual_h16res followed by dprnn_lstm_layer_bf16_pp, bit for bit, for both the LSTM output and
    the updated residual stream; tiles that straddle utterances and a ragged last tile included."""
    from tss_with_dprnn_b200.engine import Engine
    L = P.lib()
    H = 128
    dtype = torch.float16 if fmt == 'fp16' else torch.bfloat16
    h16 = 1 if fmt == 'fp16' else 0
    flags = 1 | (2 if fmt == 'fp16' else 0)
    B, S, K = (5, 60, 37) if not inter else (3, 21, 250)       # intra: 300 sequences, 60 per utterance
    rows = B * S * K
    torch.manual_seed(31 + inter)
    rnn = torch.nn.LSTM(H, H, batch_first=True, bidirectional=(ndir == 2)).cuda()
    wp2, bp = Engine._pack_lstm_tc(rnn, ['', '_reverse'][:ndir], half_jobs=True, dtype=dtype)
    x0 = rnd(rows, H, seed=1).to(dtype).cuda()
    y = (2 * rnd(rows, H, seed=2) + 0.3).to(dtype).cuda()
    mr = torch.stack([0.3 + 0.1 * rnd(B, seed=3), 0.5 + 0.1 * rnd(B, seed=4).abs()], 1).cuda().contiguous()
    gamma, beta = (1 + 0.1 * rnd(H, seed=5)).cuda(), (0.1 * rnd(H, seed=6)).cuda()
    st = stream()
    xa = x0.clone()
    L.call('dprnn_norm_residual_h16res', y, xa, None, mr, gamma, beta, B, S * K, H, h16, st)
    want = torch.empty(rows, ndir * H, device=DEV, dtype=dtype)
    L.call('dprnn_lstm_layer_bf16_pp', xa, wp2, bp, want, B, S, K, inter, H, ndir, flags, st)
    got = torch.full((rows + 1, ndir * H), 7.0, device=DEV, dtype=dtype)
    xout = torch.full((rows + 1, H), 7.0, device=DEV, dtype=dtype)
    for _ in range(2):
        L.call('dprnn_lstm_layer_bf16_pp_fused', x0, y, mr, gamma, beta, xout, wp2, bp, got, B, S, K, inter, H, ndir, flags, st)
        torch.cuda.synchronize()
        assert torch.equal(xout[:rows], xa)
        assert torch.equal(got[:rows], want)
        assert float((got[rows:].float() - 7.0).abs().max()) == 0.0 and float((xout[rows:].float() - 7.0).abs().max()) == 0.0


def test_fp16_storage_saturates_instead_of_overflowing():
    """fp16 has 5 exponent bits: the 16-bit stores of the fp16 mode clamp at +-65504 (cvt.rn.satfinite) instead of writing
    inf, and round to nearest even like torch everywhere else; bf16 is untouched."""
    x = torch.tensor([1e6, -1e6, 65504.0, 70000.0, 1.0009765625, 3.14159, -0.0, 6e-8] * 64, device=DEV)
    for h16, dtype in ((1, torch.float16), (0, torch.bfloat16)):
        out = torch.empty(x.numel(), device=DEV, dtype=dtype)
        P.lib().call('dprnn_cast_h16', x, out, x.numel(), h16, stream())
        torch.cuda.synchronize()
        want = x.to(dtype)
        if h16:
            want = torch.where(torch.isinf(want), torch.sign(want) * 65504.0, want)
        assert torch.equal(out, want), (h16, out[:8], want[:8])


@pytest.mark.parametrize('B,R,K', [(3, 1337, 256), (2, 48500, 256), (5, 700, 128), (64, 300, 256)])
@pytest.mark.parametrize('fmt', ['bf16', 'fp16'])
@pytest.mark.parametrize('discard', [0, 1, 1 | (1 << 8), 2 << 8])
def test_linear_normres_one_launch_is_bit_identical(B, R, K, fmt, discard):
    """Linear + norm + residual as ONE launch whose Linear output only lives in L2 (linear_normres.cu) == the Linear kernel
    followed by the norm kernel, bit for bit (statistics included); tiles straddle utterances, many / few utterances."""
    L = P.lib()
    M, N = B * R, 128
    dtype = torch.float16 if fmt == 'fp16' else torch.bfloat16
    h16 = 1 if fmt == 'fp16' else 0
    A = (rnd(M, K, seed=M % 1000) + 0.2).to(dtype).to(DEV)
    W = (rnd(N, K, seed=5) / K ** 0.5).to(dtype).to(DEV)
    bias, gamma, beta = rnd(N, seed=3).to(DEV), (1 + 0.1 * rnd(N, seed=4)).to(DEV), (0.1 * rnd(N, seed=6)).to(DEV)
    x0 = rnd(M, N, seed=7).to(dtype).to(DEV)
    part = torch.empty(L.query('dprnn_gemm_tc_stats_bytes', M), device=DEV, dtype=torch.uint8)
    y = torch.empty((M, N), device=DEV, dtype=dtype)
    mr = torch.empty(B, 2, device=DEV)
    L.call('dprnn_linear_h16out_stats', A, W, bias, y, M, K, part, R, 1e-5, mr, h16, stream())
    want = x0.clone()
    L.call('dprnn_norm_residual_h16res', y, want, None, mr, gamma, beta, B, R, N, h16, stream())
    ws = torch.empty(L.query('dprnn_linear_normres_workspace_bytes', B), device=DEV, dtype=torch.uint8)
    for _ in range(2):
        got = x0.clone()
        y2 = torch.full((M, N), float('nan'), device=DEV, dtype=dtype)
        mr2 = torch.full((B, 2), float('nan'), device=DEV)
        L.call('dprnn_linear_normres_h16', A, W, bias, y2, got, gamma, beta, M, K, part, R, 1e-5, mr2, ws, discard, h16, stream())
        torch.cuda.synchronize()
        assert torch.equal(mr2, mr)
        assert torch.equal(got, want)


@pytest.mark.parametrize('B,L,K,P_', [(2, 1337, 100, 50), (1, 23999, 250, 125), (3, 249, 250, 125), (2, 1, 20, 10)])
@pytest.mark.parametrize('fmt', ['bf16', 'fp16'])
def test_last_norm_residual_fold_prelu_one_pass_is_bit_identical(B, L, K, P_, fmt):
    """dprnn_norm_residual_fold_prelu_h16 (last half-block's norm + residual + PReLU + overlap-add in one pass over the
    16-bit tensors, dprnn.py:98-99,174,203-217) against the two kernels it replaces and against F.fold in fp64."""
    F = 128
    lib = P.lib()
    h16 = 1 if fmt == 'fp16' else 0
    dt = torch.float16 if h16 else torch.bfloat16
    S = lib.query('dprnn_num_chunks', L, K, P_)
    rows = B * S * K
    y = rnd(rows, F, seed=L).to(dt).to(DEV)
    xb = (0.5 * rnd(rows, F, seed=L + 1)).to(dt).to(DEV)
    mr = torch.stack([0.1 * rnd(B, seed=2), 1 + 0.2 * rnd(B, seed=3).abs()], 1).contiguous().to(DEV)
    gamma, beta = (1 + 0.1 * rnd(F, seed=4)).to(DEV), (0.1 * rnd(F, seed=6)).to(DEV)
    a = torch.tensor([0.25], device=DEV)
    xf = torch.full((rows, F), float('nan'), device=DEV)
    lib.call('dprnn_norm_residual_h16res', y, xb.clone(), xf, mr, gamma, beta, B, S * K, F, h16, stream())
    want = torch.full((B, L, F), float('nan'), device=DEV)
    lib.call('dprnn_fold_prelu', xf, want, B, L, K, P_, F, a, stream())
    got = torch.full((B, L, F), float('nan'), device=DEV)
    lib.call('dprnn_norm_residual_fold_prelu_h16', y, xb, mr, gamma, beta, got, B, L, K, P_, F, a, h16, stream())
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    # ... and the composition itself against torch in fp64
    v = xb.double().cpu() + ((y.double().cpu().view(B, -1) - mr[:, :1].double().cpu()) * mr[:, 1:].double().cpu()).view(rows, F) \
        * gamma.double().cpu() + beta.double().cpu()
    v = torch.where(v >= 0, v, 0.25 * v).view(B, S, K, F).permute(0, 3, 2, 1).reshape(B, F * K, S)
    ref = torch.nn.functional.fold(v, ((S - 1) * P_ + K, 1), kernel_size=(K, 1), stride=(P_, 1))[:, :, K:K + L, 0]
    assert O.peak_rel_err(got.cpu().double().transpose(1, 2), ref) < 1e-6
