"""Host-side logic that needs no GPU: the LSTM weight packings of the tensor-core kernels and the flat parameter buffer."""
import torch

from tss_with_dprnn_b200.engine import Engine
from tss_with_dprnn_b200.dp import FlatParams


def test_lstm_weight_packings_follow_the_documented_row_maps():
    """include/dprnn_b200.h: one-job kernel rows {q*H + 64*nh + j : q in (2r, 2r+1), j < 64}; half-job kernel rows
    {gate*H + 64*nh + 32*r + u : gate < 4, u < 32}; i, f, o rows pre-scaled by 1/2; bias [nh][gate][64 units]."""
    torch.manual_seed(0)
    H = 128
    rnn = torch.nn.LSTM(H, H, 1, batch_first=True, bidirectional=True)
    sfx = ['', '_reverse']
    wp, bp = Engine._pack_lstm_tc(rnn, sfx)
    wp2, bp2 = Engine._pack_lstm_tc(rnn, sfx, half_jobs=True)
    assert wp.shape == wp2.shape == (2 * 2 * 2 * 128, 2 * H) and wp.dtype == torch.bfloat16
    assert torch.equal(bp, bp2) and bp.shape == (2, 4 * H)
    half = torch.ones(4 * H)
    half[:2 * H] = 0.5
    half[3 * H:] = 0.5
    for d, sf in enumerate(sfx):
        wcat = torch.cat([getattr(rnn, 'weight_ih_l0' + sf), getattr(rnn, 'weight_hh_l0' + sf)], 1).detach() * half[:, None]
        b = ((getattr(rnn, 'bias_ih_l0' + sf) + getattr(rnn, 'bias_hh_l0' + sf)).detach() * half)
        for r in range(2):
            for nh in range(2):
                base = ((d * 2 + r) * 2 + nh) * 128
                rows1 = [q * H + 64 * nh + j for q in (2 * r, 2 * r + 1) for j in range(64)]
                rows2 = [g * H + 64 * nh + 32 * r + u for g in range(4) for u in range(32)]
                assert torch.equal(wp[base:base + 128], wcat[rows1].to(torch.bfloat16))
                assert torch.equal(wp2[base:base + 128], wcat[rows2].to(torch.bfloat16))
        for nh in range(2):
            for g in range(4):
                assert torch.equal(bp[d, nh * 256 + g * 64:nh * 256 + (g + 1) * 64], b[g * H + 64 * nh:g * H + 64 * nh + 64])


def test_flat_params_alignment_and_views():
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.PReLU(), torch.nn.Linear(5, 3))
    want = [p.detach().clone() for p in net.parameters()]
    fp = FlatParams(net)
    assert fp.numel == sum(p.numel() for p in want) and fp.size % FlatParams.ALIGN == 0
    for (n, p), w, off in zip(fp.named, want, fp.offsets):
        assert off % FlatParams.ALIGN == 0                          # 256-byte aligned starts (16-byte vector loads)
        assert torch.equal(p.detach(), w)                           # values preserved
        assert p.data_ptr() == fp.flat.data_ptr() + 4 * off        # the parameter aliases the flat buffer
        assert p.grad.data_ptr() == fp.grad.data_ptr() + 4 * off
    fp.grad.fill_(1.0)
    fp.zero_grad()
    assert float(fp.grad.abs().max()) == 0.0
    # padding stays zero, so norms / all-reduce / Adam over the whole buffer equal those over the parameters
    mask = torch.ones(fp.size, dtype=torch.bool)
    for (n, p), off in zip(fp.named, fp.offsets):
        mask[off:off + p.numel()] = False
    assert float(fp.flat[mask].abs().max() if mask.any() else 0.0) == 0.0
