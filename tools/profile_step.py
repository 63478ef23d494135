"""One short forward that launches every kernel family once or twice - the command profiled under ncu.

    python tools/profile_step.py [--batch 64] [--fusion cat|att|film|...] [--precision bf16|fp32]

A 1-block (n_repeats=1) DPRNN-Spe forward at the headline shape (3 s @ 8 kHz) launches the same kernels on the
same tensor sizes as the 6-block model, 1/6 as often, which keeps an `ncu --set full` capture of ALL kernels
short.  Prints the launch count so the caller can pick -s/-c.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import tss_with_dprnn_b200 as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--samples', type=int, default=24000)
    ap.add_argument('--fusion', default='cat')
    ap.add_argument('--precision', default='fp16')
    ap.add_argument('--lstm-slices', type=int, default=1, help='0 = auto: the persistent time-sliced LSTM kernel')
    ap.add_argument('--repeats', type=int, default=1)
    ap.add_argument('--passes', type=int, default=1)
    a = ap.parse_args()
    kw = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
              n_repeats=a.repeats, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0,
              fusion_type=a.fusion)
    torch.manual_seed(0)
    model = P.DPRNNSpeTasNet(**kw).eval().cuda()
    model.precision = a.precision
    model._engine.lstm_slices = a.lstm_slices
    g = torch.Generator().manual_seed(1234)
    mix = (0.05 * torch.randn(a.batch, a.samples, generator=g)).cuda()
    ref = (0.05 * torch.randn(a.batch, a.samples, generator=g)).cuda()
    rl = torch.tensor(float(a.samples))
    with torch.no_grad():
        for _ in range(a.passes):
            n0 = P.lib().launches
            est, _ = model(mix, ref, rl)
            torch.cuda.synchronize()
            print('launches per forward:', P.lib().launches - n0, ' est abs-mean', float(est.abs().mean()))


if __name__ == '__main__':
    main()
