"""The BPTT kernel alone at the cfg-5 shapes (16 x 3 s: intra 3 088 sequences x 250 steps, inter 4 000 x 193), random
saved activations - timing of the two tile sizes and a target for `ncu -k regex:lstm_bptt`.

    gpurun -- python tools/bptt_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tss_with_dprnn_b200._lib import lib  # noqa: E402

DEV = 'cuda:0'
L = lib()
B, S, K, H, nd = 16, 193, 250, 128, 2
rows = B * S * K
torch.manual_seed(0)
dh = torch.randn(rows, nd * H, device=DEV) * 0.1
gates = torch.rand(rows, nd * 4 * H, device=DEV).to(torch.bfloat16)
cst = torch.randn(rows, nd * H, device=DEV)
whhT = (torch.randn(nd, H, 4 * H, device=DEV) * 0.05).to(torch.bfloat16)
dg = torch.empty(rows, nd * 4 * H, device=DEV)
st = torch.cuda.current_stream().cuda_stream
reps = int(os.environ.get('REPS', '5'))


def timeit(fn, n=reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for inter in (0, 1):
    geo = (B * S, K, 1, K, 0, 1) if inter == 0 else (B * K, S, K, S * K, 1, K)
    for name, fl in (('64 rows per CTA', 1 | 4), ('128 rows per CTA', 1 | 8)):
        ms = timeit(lambda: L.call('dprnn_lstm_bptt_tc', dh, gates, cst, whhT, dg, *geo, H, nd, fl, st))
        gb = rows * nd * (1024 + 512 + 512 + 2048) / 1e9
        print(f'inter={inter} {name:18s} {ms:.3f} ms  {ms * 1e3 / geo[1]:.2f} us/step  {gb / ms:.0f} GB/s algorithmic')
