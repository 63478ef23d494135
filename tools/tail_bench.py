"""Micro-benchmark of the half-block tail at the headline shape (B = 64 x 3 s: 3.1 M rows): Linear kernel + norm kernel
against the one-launch form that keeps the Linear output in L2 (linear_normres.cu), without / with discarding y from L2.
CUDA events, 10 launches each after 2 warm-ups.   python tools/tail_bench.py [B]"""
import statistics, sys, torch
sys.path.insert(0, '.')
import tss_with_dprnn_b200 as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S, K, H, nd, F = 194, 250, 128, 2, 128
L = P.lib()
R = S * K
M = B * R
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
hb = torch.randn(M, nd * H, device='cuda').to(torch.bfloat16)
W = (torch.randn(F, nd * H, device='cuda') / 16).to(torch.bfloat16)
bias, gamma, beta = torch.zeros(F, device='cuda'), torch.ones(F, device='cuda'), torch.zeros(F, device='cuda')
xb = torch.randn(M, F, device='cuda').to(torch.bfloat16)
y = torch.empty(M, F, device='cuda', dtype=torch.bfloat16)
part = torch.empty(L.query('dprnn_gemm_tc_stats_bytes', M), device='cuda', dtype=torch.uint8)
mr = torch.empty(B, 2, device='cuda')
ws = torch.empty(L.query('dprnn_linear_normres_workspace_bytes', B), device='cuda', dtype=torch.uint8)


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


def two():
    L.call('dprnn_linear_h16out_stats', hb, W, bias, y, M, nd * H, part, R, 1e-5, mr, 0, st)
    L.call('dprnn_norm_residual_h16res', y, xb, None, mr, gamma, beta, B, R, F, 0, st)


gb = lambda ms, gbytes: gbytes / ms
print(f'B={B}: rows {M}; algorithmic HBM bytes: two kernels {4.76 * B / 64:.2f} GB, one launch {3.18 * B / 64:.2f} GB (+{0.79 * B / 64:.2f} if y is written back)')
lo, med = bench(two)
print(f'Linear kernel + norm kernel : min {lo:.3f} median {med:.3f} ms')
for lead in (0, 1, 2, 3, 4, 6):
    for d in (0, 1):
        lo, med = bench(lambda: L.call('dprnn_linear_normres_h16', hb, W, bias, y, xb, gamma, beta, M, nd * H, part, R, 1e-5, mr, ws,
                                       d | (lead << 8), 0, st))
        print(f'one launch, lead = {lead}, discard_y = {d}   : min {lo:.3f} median {med:.3f} ms', flush=True)
