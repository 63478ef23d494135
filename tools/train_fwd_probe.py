"""What saving the activations costs the training forward (cfg-5 shape, 16 x 3 s): the half-job kernel without saves
(inference), with the saved gates / c / h stored from the registers, and staged through shared memory + TMA.

    gpurun -- python tools/train_fwd_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tss_with_dprnn_b200._lib import lib  # noqa: E402
from tss_with_dprnn_b200.engine import Engine  # noqa: E402

DEV = 'cuda:0'
L = lib()
B, S, K, H, nd = 16, 193, 250, 128, 2
rows = B * S * K
torch.manual_seed(0)
rnn = torch.nn.LSTM(H, H, 1, batch_first=True, bidirectional=True).to(DEV)
wp, bp = Engine._pack_lstm_tc(rnn, ['', '_reverse'], half_jobs=True)
xb = (0.5 * torch.randn(rows, H, device=DEV)).to(torch.bfloat16)
hb = torch.empty(rows, nd * H, device=DEV, dtype=torch.bfloat16)
gates = torch.empty(rows, nd * 4 * H, device=DEV, dtype=torch.bfloat16)
cst = torch.empty(rows, nd * H, device=DEV)
hf = torch.empty(rows, nd * H, device=DEV)
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for inter in (0, 1):
    print(f'inter={inter}: inference kernel (no saves) '
          f"{timeit(lambda: L.call('dprnn_lstm_layer_bf16_pp', xb, wp, bp, hb, B, S, K, inter, H, nd, 1, st)):.3f} ms")
    for name, fl in (('128-seq tiles, direct stores', 1 | 16), ('128-seq tiles, staged + TMA', 1), ('256-seq tiles, direct', 1 | 8)):
        ms = timeit(lambda: L.call('dprnn_lstm_layer_bf16_train_pp', xb, wp, bp, hb, gates, cst, hf, B, S, K, inter, H, nd, fl, st))
        print(f'  {name:28s} {ms:.3f} ms')
