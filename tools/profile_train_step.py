"""One short TRAINING iteration (cfg 5 shape: DPRNN-Spe FiLM, 16 x 3 s @ 8 kHz, tensor-core mode) of a 1-block model -
the command profiled under ncu for the backward / weight-gradient kernels (same tensor sizes as the 6-block model,
1/6 of the launches).

    python tools/profile_train_step.py [--batch 16] [--precision bf16|fp32] [--repeats 1] [--iters 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import tss_with_dprnn_b200 as P  # noqa: E402
from tss_with_dprnn_b200.train import SpeTrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--samples', type=int, default=24000)
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--repeats', type=int, default=1)
    ap.add_argument('--iters', type=int, default=2)
    a = ap.parse_args()
    kw = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
              n_repeats=a.repeats, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0,
              fusion_type='film')
    torch.manual_seed(0)
    model = P.DPRNNSpeTasNet(**kw).cuda()
    model.precision = a.precision
    stepper = SpeTrainStep(model)
    g = torch.Generator().manual_seed(1234)
    mix, ref, tgt = ((0.05 * torch.randn(a.batch, a.samples, generator=g)).cuda() for _ in range(3))
    spk = torch.randint(0, 251, (a.batch,), generator=g).cuda()
    for _ in range(a.iters):
        n0 = P.lib().launches
        loss = stepper.step(mix, ref, tgt, spk, ref_len=a.samples)
        torch.cuda.synchronize()
        print('launches per iteration:', P.lib().launches - n0, ' loss', [round(float(v), 4) for v in loss])


if __name__ == '__main__':
    main()
