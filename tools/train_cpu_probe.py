"""Is the training step launch-bound?  CPU time to ISSUE one step (no synchronisation) next to the GPU time of the step.

    gpurun -- python tools/train_cpu_probe.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tss_with_dprnn_b200 as P  # noqa: E402
from tss_with_dprnn_b200.train import SpeTrainStep  # noqa: E402

kw = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125, n_repeats=6,
          bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0, fusion_type='film')
torch.manual_seed(0)
model = P.DPRNNSpeTasNet(**kw).cuda().train()
model.precision = 'bf16'
step = SpeTrainStep(model)
B, T = 16, 24000
g = torch.Generator(device='cuda').manual_seed(1)
mix, ref, tgt = (0.05 * torch.randn(B, T, device='cuda', generator=g) for _ in range(3))
spk = torch.randint(0, 251, (B,), device='cuda')
for _ in range(3):
    step.step(mix, ref, tgt, spk)
torch.cuda.synchronize()
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    step.step(mix, ref, tgt, spk)
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'CPU issue time {1e3 * (t1 - t0):.1f} ms, GPU step {a.elapsed_time(b):.1f} ms, wall incl. sync {1e3 * (t2 - t0):.1f} ms')
