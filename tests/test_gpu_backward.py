"""Backward building blocks (cfg 5) against torch autograd / fp64 restatements."""
import pytest
import torch

import tss_with_dprnn_b200 as P

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def st():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def pack_gates(g, nd, H=128):
    """fp32 row-major gates [rows, nd*4H] -> the bf16 layout of the tensor-core kernels [rows][nd][16][4][8]."""
    rows = g.shape[0]
    return g.view(rows, nd, 4, H // 8, 8).permute(0, 1, 3, 2, 4).contiguous().to(torch.bfloat16).view(rows, nd * 4 * H)


def unpack_gates(gp, nd, H=128):
    rows = gp.shape[0]
    return gp.view(rows, nd, H // 8, 4, 8).permute(0, 1, 3, 2, 4).contiguous().float().view(rows, nd * 4 * H)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('M,N1,N2', [(1000, 128, 64), (48500, 512, 128), (20001, 256, 128), (33, 64, 64)])
def test_gemm_atb(M, N1, N2):
    L = P.lib()
    A, B = rnd(M, N1, seed=1), rnd(M, N2, seed=2)
    want = A.double().t() @ B.double()
    C = torch.full((N1, N2), 3.0, device=DEV)
    ws = torch.empty(L.query('dprnn_gemm_atb_workspace_bytes', M, N1, N2), device=DEV, dtype=torch.uint8)
    L.call('dprnn_gemm_atb', A.to(DEV), N1, B.to(DEV), N2, C, N2, M, N1, N2, 0, ws, st())
    assert rel(C.cpu(), want) < 1e-5
    L.call('dprnn_gemm_atb', A.to(DEV), N1, B.to(DEV), N2, C, N2, M, N1, N2, 1, ws, st())
    assert rel(C.cpu(), 2 * want) < 1e-5
    out = torch.zeros(N1, device=DEV)
    ws2 = torch.empty(L.query('dprnn_col_sum_workspace_bytes', N1), device=DEV, dtype=torch.uint8)
    L.call('dprnn_col_sum', A.to(DEV), N1, None, 0, M, N1, out, 0, ws2, st())
    assert rel(out.cpu(), A.double().sum(0)) < 1e-5


@pytest.mark.parametrize('M,N1,N2,lda,ldb', [(50017, 512, 128, 1024, 256), (48500, 128, 256, 128, 256),
                                             (20000, 128, 128, 128, 128), (4100, 256, 128, 256, 128),
                                             (100, 512, 128, 512, 128), (30011, 64, 128, 64, 128),
                                             (30011, 128, 64, 128, 64)])
def test_gemm_atb_tensor_core(M, N1, N2, lda, ldb):
    """The MN-major TF32 weight-gradient contraction against fp64 (operands are column slices of wider row-major
    tensors, as the per-direction slices of dgates / h_prev are)."""
    L = P.lib()
    assert L.query('dprnn_gemm_atb_tc_supported', N1, N2, lda, ldb)
    Af, Bf = rnd(M, lda, seed=1).to(DEV), rnd(M, ldb, seed=2).to(DEV)
    a_off, b_off = lda - N1, ldb - N2                    # the last N columns of each
    want = Af[:, a_off:].double().t() @ Bf[:, b_off:].double()
    C = torch.full((N1, N2 + 4), 3.0, device=DEV)       # ldc > N2: the padding must stay untouched
    ws = torch.empty(L.query('dprnn_gemm_atb_tc_workspace_bytes', N1, N2), device=DEV, dtype=torch.uint8)
    L.call('dprnn_gemm_atb_tc', Af.data_ptr() + 4 * a_off, lda, Bf.data_ptr() + 4 * b_off, ldb, C, N2 + 4, M, N1, N2, 0,
           ws, st())
    scale = float(want.abs().max())
    assert float((C[:, :N2].double() - want).abs().max()) < 3e-3 * scale
    assert float((C[:, N2:] - 3.0).abs().max()) == 0.0
    first = C.clone()
    L.call('dprnn_gemm_atb_tc', Af.data_ptr() + 4 * a_off, lda, Bf.data_ptr() + 4 * b_off, ldb, C, N2 + 4, M, N1, N2, 1,
           ws, st())
    assert torch.equal(C[:, :N2], 2 * first[:, :N2])     # deterministic reduction order, accumulate adds exactly


def test_groupnorm_bwd():
    L = P.lib()
    B, R, C = 3, 777, 128
    y = rnd(B, R, C, seed=3).double().requires_grad_(True)
    gamma = (1 + 0.1 * rnd(C, seed=4)).double().requires_grad_(True)
    beta = rnd(C, seed=5).double().requires_grad_(True)
    mean = y.mean((1, 2), keepdim=True); var = y.var((1, 2), unbiased=False, keepdim=True)
    z = (y - mean) / torch.sqrt(var + 1e-5) * gamma + beta
    dz = rnd(B, R, C, seed=6).double()
    z.backward(dz)
    mr = torch.stack([mean.flatten(), 1 / torch.sqrt(var.flatten() + 1e-5)], 1).float().to(DEV)
    dy = torch.empty(B, R, C, device=DEV); dg = torch.zeros(C, device=DEV); db = torch.zeros(C, device=DEV)
    ws = torch.empty(L.query('dprnn_gn_bwd_workspace_bytes', B, C), device=DEV, dtype=torch.uint8)
    L.call('dprnn_groupnorm_bwd', dz.float().to(DEV), y.detach().float().to(DEV), mr, gamma.detach().float().to(DEV), B, R, C,
           dy, 0, dg, db, ws, st())
    assert rel(dy.cpu(), y.grad) < 2e-5
    assert rel(dg.cpu(), gamma.grad) < 2e-5 and rel(db.cpu(), beta.grad) < 2e-5


@pytest.mark.parametrize('inter', [0, 1])
@pytest.mark.parametrize('ndir', [2, 1])
def test_lstm_train_forward_and_bptt(inter, ndir, big=False):
    """dprnn_lstm_recurrence_f32_train + dprnn_lstm_bptt_f32 against autograd through nn.LSTM (fp64)."""
    L = P.lib()
    H = 128
    B, S, K = (2, 5, 11) if not inter else (2, 7, 9)          # 10 / 18 sequences: one padded 256-sequence tile
    if big:
        B, S, K = 3, 50, 100                                  # 150 / 300 sequences: both CTAs of a pair, two pair-jobs
    torch.manual_seed(7 + inter)
    rnn = torch.nn.LSTM(H, H, batch_first=True, bidirectional=(ndir == 2)).double()
    x = rnd(B, S, K, H, seed=8).double().requires_grad_(True)
    seqs = x.reshape(B * S, K, H) if not inter else x.permute(0, 2, 1, 3).reshape(B * K, S, H)
    out, _ = rnn(seqs)
    out_l = out.reshape(B, S, K, ndir * H) if not inter else out.reshape(B, K, S, ndir * H).permute(0, 2, 1, 3)
    dout = rnd(B, S, K, ndir * H, seed=9).double()
    out_l.backward(dout)
    rows = B * S * K
    sfx = ['', '_reverse'][:ndir]
    wih = torch.cat([getattr(rnn, 'weight_ih_l0' + s) for s in sfx], 0).detach().float()       # [nd*4H, H]
    bias = torch.cat([getattr(rnn, 'bias_ih_l0' + s) + getattr(rnn, 'bias_hh_l0' + s) for s in sfx], 0).detach().float()
    whh = torch.stack([getattr(rnn, 'weight_hh_l0' + s) for s in sfx], 0).detach().float()      # [nd, 4H, H]
    gx = (x.detach().float().reshape(rows, H) @ wih.t() + bias).to(DEV)
    geo = (B * S, K, 1, K, 0, 1) if not inter else (B * K, S, K, S * K, 1, K)
    hout = torch.empty(rows, ndir * H, device=DEV); gates = torch.empty(rows, ndir * 4 * H, device=DEV)
    cst = torch.empty(rows, ndir * H, device=DEV)
    L.call('dprnn_lstm_recurrence_f32_train', gx, whh.transpose(1, 2).contiguous().to(DEV), hout, gates, cst, *geo, H, ndir, st())
    assert rel(hout.cpu(), out_l.detach().reshape(rows, -1)) < 1e-5
    dg = torch.empty(rows, ndir * 4 * H, device=DEV)
    L.call('dprnn_lstm_bptt_f32', dout.float().reshape(rows, -1).contiguous().to(DEV), gates, cst, whh.contiguous().to(DEV), dg,
           *geo, H, ndir, st())
    torch.cuda.synchronize()
    dgc = dg.cpu().double()
    # the tensor-core BPTT (bf16 recurrent contraction) from the same saved activations: within bf16 tolerance of the exact
    # kernel; rows past the real ones are never written
    dg2 = torch.full((rows + 1, ndir * 4 * H), 7.0, device=DEV)
    whhT = whh.transpose(1, 2).contiguous().to(torch.bfloat16).to(DEV)
    gates_p = pack_gates(gates, ndir)
    # flags: tanh.approx (1) | 64 rows per CTA (4) | 128 rows per CTA (8); 0 / 1 leave the tile size to the library
    got = {}
    for flags in (0, 1, 4, 5, 8, 9):
        dg2[:rows].zero_()
        L.call('dprnn_lstm_bptt_tc', dout.float().reshape(rows, -1).contiguous().to(DEV), gates_p, cst, whhT, dg2, *geo, H,
               ndir, flags, st())
        assert float((dg2[:rows] - dg).abs().max()) < 2e-2 * float(dg.abs().max()), flags
        assert float((dg2[rows:] - 7.0).abs().max()) == 0.0
        got[flags] = dg2[:rows].clone()
    # both tile sizes run the same arithmetic per row (M = 128 and M = 256 MMAs accumulate the K = 512 in the same order)
    assert torch.equal(got[4], got[8]) and torch.equal(got[5], got[9])
    assert torch.equal(got[0], got[4]) or torch.equal(got[0], got[8])
    # bf16 output (TMA stores of the tensor core's operand tile): the bf16 rounding of the fp32 output, bit for bit
    for flags in (5, 9):
        dgb = torch.full((rows + 1, ndir * 4 * H), 7.0, device=DEV, dtype=torch.bfloat16)
        L.call('dprnn_lstm_bptt_tc_bf16out', dout.float().reshape(rows, -1).contiguous().to(DEV), gates_p, cst, whhT, dgb, *geo,
               H, ndir, flags, st())
        assert torch.equal(dgb[:rows], got[flags].to(torch.bfloat16)), flags
        assert float((dgb[rows:].float() - 7.0).abs().max()) == 0.0
    # dx = dgates @ W_ih ; dW_ih = dgates^T x ; db = colsum(dgates)
    dx = dgc @ wih.double()
    assert rel(dx, x.grad.reshape(rows, H)) < 1e-4
    dwih = dgc.t() @ x.detach().reshape(rows, H)
    want_wih = torch.cat([getattr(rnn, 'weight_ih_l0' + s).grad for s in sfx], 0)
    assert rel(dwih, want_wih) < 1e-4
    want_b = torch.cat([getattr(rnn, 'bias_ih_l0' + s).grad for s in sfx], 0)
    assert rel(dgc.sum(0), want_b) < 1e-4


@pytest.mark.parametrize('inter', [0, 1])
def test_lstm_bptt_many_sequences(inter):
    test_lstm_train_forward_and_bptt(inter, 2, big=True)


@pytest.mark.parametrize('inter', [0, 1])
def test_lstm_tensor_core_train_forward_saves_what_bptt_needs(inter):
    """dprnn_lstm_layer_bf16_train against the exact fp32 training recurrence: gate activations, cell state and h of
    every step land at the same [rows, .] positions (bf16-operand tolerance), padded tile rows are never written."""
    from tss_with_dprnn_b200.engine import Engine
    L = P.lib()
    B, S, K, H, F, nd = 2, 5, 250, 128, 128, 2
    torch.manual_seed(3)
    rnn = torch.nn.LSTM(F, H, 1, batch_first=True, bidirectional=True).to(DEV)
    rows = B * S * K
    xs = (0.5 * rnd(rows, F, seed=4)).to(DEV)
    sfx = ['', '_reverse']
    wih = torch.cat([getattr(rnn, 'weight_ih_l0' + s).detach() for s in sfx], 0)
    b = torch.cat([(getattr(rnn, 'bias_ih_l0' + s) + getattr(rnn, 'bias_hh_l0' + s)).detach() for s in sfx], 0)
    whh = torch.stack([getattr(rnn, 'weight_hh_l0' + s).detach() for s in sfx], 0).contiguous()
    gx = xs @ wih.t() + b
    geo = (B * S, K, 1, K, 0, 1) if inter == 0 else (B * K, S, K, S * K, 1, K)
    h0, g0, c0 = (torch.empty(rows, n, device=DEV) for n in (nd * H, nd * 4 * H, nd * H))
    L.call('dprnn_lstm_recurrence_f32_train', gx.contiguous(), whh.transpose(1, 2).contiguous(), h0, g0, c0, *geo, H, nd, st())
    xb = xs.to(torch.bfloat16)
    wp, bp = Engine._pack_lstm_tc(rnn, sfx)
    hb = torch.empty(rows, nd * H, device=DEV, dtype=torch.bfloat16)
    h1, c1 = (torch.full((rows + 1, n), 7.0, device=DEV) for n in (nd * H, nd * H))
    g1p = torch.full((rows + 1, nd * 4 * H), 7.0, device=DEV, dtype=torch.bfloat16)
    wp2, _ = Engine._pack_lstm_tc(rnn, sfx, half_jobs=True)
    # flag word of the half-job kernel: 1 tanh.approx | 4 / 8 = 128- / 256-sequence tiles | 16 = the 128-sequence tiles store
    # the saved values from the registers instead of through staging buffers + TMA
    saved = {}
    for fast, fn, w in ((0, 'dprnn_lstm_layer_bf16_train', wp), (1, 'dprnn_lstm_layer_bf16_train', wp),
                        (1, 'dprnn_lstm_layer_bf16_train_pp', wp2), (1 | 4, 'dprnn_lstm_layer_bf16_train_pp', wp2),
                        (1 | 4 | 16, 'dprnn_lstm_layer_bf16_train_pp', wp2), (1 | 8, 'dprnn_lstm_layer_bf16_train_pp', wp2),
                        (4, 'dprnn_lstm_layer_bf16_train_pp', wp2), (4 | 16, 'dprnn_lstm_layer_bf16_train_pp', wp2)):
        g1p.fill_(7.0); c1.fill_(7.0); h1.fill_(7.0)
        L.call(fn, xb, w, bp, hb, g1p, c1, h1, B, S, K, inter, H, nd, fast, st())
        g1 = unpack_gates(g1p, nd)
        assert float((g1[:rows] - g0).abs().max()) < 3e-2, (fast, fn)
        assert float((c1[:rows] - c0).abs().max()) < 5e-2, (fast, fn)
        assert float((h1[:rows] - h0).abs().max()) < 3e-2, (fast, fn)
        assert float((hb.float() - h1[:rows]).abs().max()) < 1e-2          # the bf16 copy of the same h
        for t_ in (g1, c1, h1):
            assert float((t_[rows:] - 7.0).abs().max()) == 0.0             # nothing written past the real rows
        if fn.endswith('_pp'):
            saved[fast] = (g1p.clone(), c1.clone(), h1.clone(), hb.clone())
    # without the fp32 copy of h (hout_f32 = NULL: h is kept as bf16 only) the other outputs do not change
    for fl in (1 | 4, 1 | 4 | 16, 1 | 8):
        g1p.fill_(7.0); c1.fill_(7.0); hb.fill_(7.0)
        L.call('dprnn_lstm_layer_bf16_train_pp', xb, wp2, bp, hb, g1p, c1, None, B, S, K, inter, H, nd, fl, st())
        assert all(torch.equal(x_, y_) for x_, y_ in zip((g1p, c1, hb), (saved[fl][0], saved[fl][1], saved[fl][3]))), fl
    # the staged epilogue stores exactly what the direct one stores; tile size does not change a row's arithmetic
    for a, b_ in ((1 | 4, 1 | 4 | 16), (1 | 4, 1 | 8), (4, 4 | 16)):
        assert all(torch.equal(x_, y_) for x_, y_ in zip(saved[a], saved[b_])), (a, b_)


@pytest.mark.parametrize('M,N1,lda', [(48500 * 4, 512, 1024), (20001, 128, 128), (776000, 512, 1024), (4100, 256, 256)])
def test_gemm_atb_tc_colsum(M, N1, lda):
    """dW = A^T B with the column sums of A (the bias gradient) from the SAME pass: a 32-column group of ones extends B in
    shared memory.  TF32 operands (truncated) against the fp64 product; strided A (one direction of a [rows, ndir*4H]
    d-gates tensor); accumulate into C, overwrite the column sums; deterministic."""
    L = P.lib()
    g = torch.Generator().manual_seed(M % 977)
    Afull = torch.randn(M, lda, generator=g)
    B = torch.randn(M, 128, generator=g)
    A = Afull[:, lda - N1:]                                      # the last N1 columns: a strided view
    tr = lambda t: (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32).double()
    want = tr(A).t() @ tr(B)
    want_cs = tr(A).sum(0)
    Ad, Bd = Afull.to(DEV), B.to(DEV)
    aptr = Ad.data_ptr() + 4 * (lda - N1)
    ws = torch.empty(L.query('dprnn_gemm_atb_tc_colsum_workspace_bytes', N1), device=DEV, dtype=torch.uint8)
    outs = []
    for _ in range(2):
        C = torch.full((N1, 128), 3.0, device=DEV)
        cs = torch.full((N1,), 7.0, device=DEV)
        L.call('dprnn_gemm_atb_tc_colsum', aptr, lda, Bd, 128, C, 128, cs, M, N1, 128, 1, 0, ws, st())
        torch.cuda.synchronize()
        outs.append((C.cpu(), cs.cpu()))
    C, cs = outs[0]
    tol = 2e-5 if M < 300000 else 1e-4           # fp32 accumulation over M rows (a split holds ~M / 37 of them)
    assert rel(C - 3.0, want) < tol
    assert float((cs.double() - want_cs).abs().max() / want_cs.abs().max()) < tol
    assert torch.equal(outs[1][0], C) and torch.equal(outs[1][1], cs)


@pytest.mark.parametrize('bf', [0, 1])
@pytest.mark.parametrize('M,K,lda', [(48500 * 4, 1024, 1024), (1000, 512, 512), (257, 1024, 1536), (776000, 1024, 1024), (300, 64, 64)])
def test_gemm_kdeep(M, K, lda, bf):
    """C[M,128] (+)= A[M,K] W[128,K]^T with W streamed next to A (256-row tiles, TMA reduce-add into C): TF32 operands
    (truncated) or bf16 operands against the fp64 product of the same operands; strided A, a ragged last tile, rows past M
    untouched, deterministic."""
    L = P.lib()
    g = torch.Generator().manual_seed(M % 977 + K)
    if bf:
        tr = lambda t: t.to(torch.bfloat16).double()
        dt, el = torch.bfloat16, 2
    else:
        tr = lambda t: (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32).double()
        dt, el = torch.float32, 4
    Afull = torch.randn(M, lda, generator=g)
    W = torch.randn(128, K, generator=g) / K ** 0.5
    C0 = torch.randn(M + 3, 128, generator=g)
    A = Afull[:, lda - K:]
    want = tr(A) @ tr(W).t()
    Ad, Wd = Afull.to(dt).to(DEV), W.to(dt).to(DEV)
    aptr = Ad.data_ptr() + el * (lda - K)
    ws = torch.empty(L.query('dprnn_gemm_kdeep_workspace_bytes'), device=DEV, dtype=torch.uint8)
    assert L.query('dprnn_gemm_kdeep_supported', bf, 128, K, lda, 128) == 1
    outs = []
    for acc in (0, 1, 1):
        C = C0.to(DEV)
        L.call('dprnn_gemm_kdeep', aptr, bf, lda, Wd, C, 128, M, 128, K, acc, ws, st())
        torch.cuda.synchronize()
        ref = want + (C0[:M].double() if acc else 0.0)
        err = float((C[:M].cpu().double() - ref).abs().max()) / float(ref.abs().max())
        assert err < 1e-4, (acc, err)
        assert torch.equal(C[M:].cpu(), C0[M:])                         # rows past M are not touched
        outs.append(C.clone())
    assert torch.equal(outs[1], outs[2])
    assert L.query('dprnn_gemm_kdeep_supported', bf, 64, K, lda, 128) == 0


@pytest.mark.parametrize('bf', [0, 1])
@pytest.mark.parametrize('B,S,K,inter,N1,shift', [(2, 37, 50, 0, 512, -1), (2, 37, 50, 0, 512, 1), (3, 41, 70, 1, 512, -1),
                                                  (3, 41, 70, 1, 512, 1), (2, 33, 250, 0, 128, 0), (16, 97, 250, 1, 512, -1)])
def test_gemm_atb_dual(B, S, K, inter, N1, shift, bf):
    """dW_ih, dW_hh and db of one LSTM direction from one pass over its d gates: C1 += A^T B1, C2 += A^T shift_t(B2),
    colsum = sum A, with B2 read one time step earlier / later through the tensor map (zero outside the sequence).  TF32
    (truncated) or bf16 operands against the fp64 products of the same operands; strided operands; accumulate into C,
    overwrite the sums."""
    L = P.lib()
    rows = B * S * K
    g = torch.Generator().manual_seed(rows % 977 + N1 + shift)
    if bf:
        tr = lambda t: t.to(torch.bfloat16).double()
        dt, el = torch.bfloat16, 2
    else:
        tr = lambda t: (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32).double()
        dt, el = torch.float32, 4
    lda, ldb2 = 2 * N1, 256
    Afull = torch.randn(rows, lda, generator=g)
    B1 = torch.randn(rows, 128, generator=g)
    B2full = torch.randn(rows, ldb2, generator=g)
    A, B2 = Afull[:, N1:], B2full[:, 128:]
    # the time shift in the [B, S, K] layout: intra -> along K, inter -> along S
    h = B2.reshape(B, S, K, 128)
    hs = torch.zeros_like(h)
    if shift == 0:
        hs = h
    elif inter == 0:
        if shift < 0: hs[:, :, 1:] = h[:, :, :-1]
        else: hs[:, :, :-1] = h[:, :, 1:]
    else:
        if shift < 0: hs[:, 1:] = h[:, :-1]
        else: hs[:, :-1] = h[:, 1:]
    want1 = tr(A).t() @ tr(B1)
    want2 = tr(A).t() @ tr(hs.reshape(rows, 128))
    want_cs = tr(A).sum(0)
    Ad, B1d, B2d = Afull.to(dt).to(DEV), B1.to(dt).to(DEV), B2full.to(dt).to(DEV)
    assert L.query('dprnn_gemm_atb_dual_supported', bf, N1, lda, 128, ldb2) == 1
    ws = torch.empty(L.query('dprnn_gemm_atb_dual_workspace_bytes', N1), device=DEV, dtype=torch.uint8)
    outs = []
    for _ in range(2):
        C1 = torch.full((N1, 128), 3.0, device=DEV)
        C2 = torch.full((N1, 256), -2.0, device=DEV)              # written at columns 128.. with ldc = 256
        cs = torch.full((N1,), 9.0, device=DEV)
        L.call('dprnn_gemm_atb_dual', Ad.data_ptr() + el * N1, bf, lda, N1, B1d, 128, B2d.data_ptr() + el * 128, ldb2, B, S, K,
               inter, shift, C1, 128, C2.data_ptr() + 4 * 128, 256, cs, 1, 0, ws, st())
        torch.cuda.synchronize()
        tol = 1e-4 * float(want1.abs().max())
        assert float((C1.cpu().double() - 3.0 - want1).abs().max()) < tol
        assert float((C2[:, 128:].cpu().double() + 2.0 - want2).abs().max()) < tol
        assert float((C2[:, :128] + 2.0).abs().max()) == 0.0
        assert float((cs.cpu().double() - want_cs).abs().max()) < 1e-4 * float(want_cs.abs().max()) + 1e-3
        outs.append((C1.clone(), C2.clone(), cs.clone()))
    assert all(torch.equal(a, b) for a, b in zip(*outs))


@pytest.mark.parametrize('rows,C', [(2 * 1300 - 1, 128), (384001, 256), (7, 64), (100003, 64)])
def test_batchnorm_training_statistics(rows, C):
    """dprnn_batchnorm_affine in training mode: per-channel batch statistics (float4 loads, fp32 over 16 rows then fp64,
    fixed partition) -> scale / shift and the running-statistics update of nn.BatchNorm1d."""
    L = P.lib()
    g = torch.Generator().manual_seed(rows + C)
    y = (torch.randn(rows, C, generator=g) * 1.7 + 0.3 * torch.arange(C)).float()
    bn = torch.nn.BatchNorm1d(C).double()
    with torch.no_grad():
        bn.weight.copy_(torch.randn(C, generator=g)); bn.bias.copy_(torch.randn(C, generator=g))
        bn.running_mean.copy_(torch.randn(C, generator=g)); bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    rm, rv = bn.running_mean.clone().float().to(DEV), bn.running_var.clone().float().to(DEV)
    want = bn.train()(y.double().t().unsqueeze(0)).squeeze(0).t().detach()         # [rows, C]
    ws = torch.empty(L.query('dprnn_bn_workspace_bytes', C), device=DEV, dtype=torch.uint8)
    scale, shift = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    yd = y.to(DEV)
    L.call('dprnn_batchnorm_affine', yd, rows, C, bn.weight.detach().float().to(DEV), bn.bias.detach().float().to(DEV), rm, rv, 1,
           float(bn.eps), float(bn.momentum), ws, scale, shift, st())
    got = yd.double() * scale.double() + shift.double()
    assert float((got.cpu() - want).abs().max()) < 2e-5 * float(want.abs().max())
    assert float((rm.cpu().double() - bn.running_mean).abs().max()) < 1e-6 * float(bn.running_mean.abs().max()) + 1e-6
    assert float((rv.cpu().double() - bn.running_var).abs().max()) < 1e-5 * float(bn.running_var.abs().max())
