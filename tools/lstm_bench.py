"""Micro-benchmark of the tensor-core LSTM layer kernels at the headline shape (B = 64 x 3 s: intra 12416 x 250, inter
16000 x 194): one-job-per-pair half-job kernel vs the persistent time-sliced one for k = 1..8.  CUDA events, 10 launches
each after 2 warm-ups, min / median.   python tools/lstm_bench.py [B]"""
import statistics, sys, torch
sys.path.insert(0, '.')
import tss_with_dprnn_b200 as P
from tss_with_dprnn_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S, K, H, nd = 194, 250, 128, 2
L = P.lib()
torch.manual_seed(0)
rnn = torch.nn.LSTM(H, H, 1, batch_first=True, bidirectional=True).cuda()
rows = B * S * K
xb = (0.5 * torch.randn(rows, H, device='cuda')).to(torch.bfloat16)
wp2, bp = Engine._pack_lstm_tc(rnn, ['', '_reverse'], half_jobs=True)
hb = torch.empty(rows, nd * H, device='cuda', dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream


def bench(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), statistics.median(ts)


for inter in (0, 1):
    J = (B if inter else (B * S + 255) // 256) * nd
    T = S if inter else K
    lo, med = bench(lambda: L.call('dprnn_lstm_layer_bf16_pp', xb, wp2, bp, hb, B, S, K, inter, H, nd, 1, st))
    print(f"{'inter' if inter else 'intra'} J={J} T={T}  pp (one job per pair): min {lo:.3f} median {med:.3f} ms   "
          f"ideal packing x{J / 74 / -(-J // 74):.3f}", flush=True)
    y = torch.randn(rows, H, device='cuda').to(torch.bfloat16)
    mr = torch.stack([torch.zeros(B), torch.ones(B)], 1).cuda().contiguous()
    gamma, beta = torch.ones(H, device='cuda'), torch.zeros(H, device='cuda')
    xo = torch.empty_like(xb)
    lo, med = bench(lambda: L.call('dprnn_lstm_layer_bf16_pp_fused', xb, y, mr, gamma, beta, xo, wp2, bp, hb, B, S, K, inter, H, nd, 1, st))
    print(f"   pp + fused input norm: min {lo:.3f} median {med:.3f} ms   (stand-alone norm pass: ", end='')
    lo, med = bench(lambda: L.call('dprnn_norm_residual_h16res', y, xo, None, mr, gamma, beta, B, S * K, H, 0, st))
    print(f"min {lo:.3f} median {med:.3f} ms)", flush=True)
    ws = torch.empty(L.query('dprnn_lstm_sliced_workspace_bytes', B, S, K, inter, nd), device='cuda', dtype=torch.uint8)
    for k in (3, 4):
        lo, med = bench(lambda: L.call('dprnn_lstm_layer_bf16_sliced', xb, wp2, bp, hb, B, S, K, inter, H, nd, 1, k, 0, ws, st))
        rounds = -(-k * J // 74)
        print(f"   sliced k={k}: min {lo:.3f} median {med:.3f} ms   ({rounds} rounds x {-(-T // k)} steps = {rounds * -(-T // k)} step-times)", flush=True)
