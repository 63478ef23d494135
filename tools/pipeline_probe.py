"""Feasibility probe for SM partitioning: the 12 half-blocks (LSTM -> Linear -> norm + residual) of the
headline batch as G utterance groups on G streams, with the LSTM kernels of all groups chained by events (never two at
once) and capped at P resident CTA pairs, so that one group's HBM-bound tail runs on the SMs the other group's LSTM
leaves free.  Prints ms per 12 half-blocks for a few (G, P, k).   python tools/pipeline_probe.py"""
import sys, torch
sys.path.insert(0, '.')
import tss_with_dprnn_b200 as P_
from tss_with_dprnn_b200.engine import Engine
L = P_.lib()
B, S, K, H, nd, F = 64, 194, 250, 128, 2, 128
torch.manual_seed(0)
rnn = torch.nn.LSTM(H, H, 1, batch_first=True, bidirectional=True).cuda()
wp2, bp = Engine._pack_lstm_tc(rnn, ['', '_reverse'], half_jobs=True)
lin_w = (torch.randn(F, nd * H, device='cuda') / 16).to(torch.bfloat16)
lin_b = torch.zeros(F, device='cuda')
gamma, beta = torch.ones(F, device='cuda'), torch.zeros(F, device='cuda')


class Group:
    def __init__(self, b):
        self.b, self.rows = b, b * S * K
        self.xb = (0.5 * torch.randn(self.rows, H, device='cuda')).to(torch.bfloat16)
        self.hb = torch.empty(self.rows, nd * H, device='cuda', dtype=torch.bfloat16)
        self.y = torch.empty(self.rows, F, device='cuda', dtype=torch.bfloat16)
        self.part = torch.empty(L.query('dprnn_gemm_tc_stats_bytes', self.rows), device='cuda', dtype=torch.uint8)
        self.mr = torch.empty(b, 2, device='cuda')
        self.ws = [torch.empty(L.query('dprnn_lstm_sliced_workspace_bytes', b, S, K, i, nd), device='cuda', dtype=torch.uint8) for i in (0, 1)]
        self.stream = torch.cuda.Stream()

    def lstm(self, inter, k, pairs):
        st = torch.cuda.current_stream().cuda_stream
        if k == 1 and pairs == 0:
            L.call('dprnn_lstm_layer_bf16_pp', self.xb, wp2, bp, self.hb, self.b, S, K, inter, H, nd, 1, st)
        else:
            L.call('dprnn_lstm_layer_bf16_sliced', self.xb, wp2, bp, self.hb, self.b, S, K, inter, H, nd, 1, k, pairs, self.ws[inter], st)

    def tail(self):
        st = torch.cuda.current_stream().cuda_stream
        L.call('dprnn_linear_h16out_stats', self.hb, lin_w, lin_b, self.y, self.rows, nd * H, self.part, S * K, 1e-5, self.mr, 0, st)
        L.call('dprnn_norm_residual_h16res', self.y, self.xb, None, self.mr, gamma, beta, self.b, S * K, F, 0, st)


def run(groups, k, pairs, chain):
    main = torch.cuda.current_stream()
    ready = torch.cuda.Event(); ready.record(main)
    for g in groups:
        g.stream.wait_event(ready)
    last = None
    for hb in range(12):
        for g in groups:
            with torch.cuda.stream(g.stream):
                if chain and last is not None:
                    g.stream.wait_event(last)
                g.lstm(hb & 1, k, pairs)
                if chain:
                    last = torch.cuda.Event(); last.record(g.stream)
                g.tail()
    for g in groups:
        ev = torch.cuda.Event(); ev.record(g.stream); main.wait_event(ev)


def bench(G, k, pairs, chain, n=5):
    base, extra = divmod(B, G)
    groups = [Group(base + (1 if i < extra else 0)) for i in range(G)]
    for _ in range(2):
        run(groups, k, pairs, chain)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run(groups, k, pairs, chain)
    e1.record()
    torch.cuda.synchronize()
    print(f'groups {G}  slices {k}  lstm pairs {pairs or 74:2d}  chained {int(chain)}: {e0.elapsed_time(e1) / n:7.2f} ms per 12 half-blocks', flush=True)
    del groups
    torch.cuda.empty_cache()


bench(1, 1, 0, False)            # today, one stream
bench(3, 1, 0, False)            # today's default: 3 unsynchronised groups
bench(1, 0, 0, False)            # sliced, one stream
for P in (40, 44, 48, 52, 56, 60):
    bench(2, 0, P, True)
for P in (44, 48, 52):
    bench(3, 0, P, True)
bench(2, 0, 48, False)
