"""Per-launch timeline of one forward (CUDA events around every C-ABI launch, on the stream it was issued to):
which kernels overlap across streams, where the gaps are.   python tools/timeline.py [--streams N]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tss_with_dprnn_b200 as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--streams', type=int, default=1)
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--rows', type=int, default=80)
    a = ap.parse_args()
    kw = dict(input_size=64, feature_size=128, hidden_size=128, chunk_length=250, kernel_size=2, hop_length=125,
              n_repeats=6, bidirectional=True, norm_type='ln', activation_type='sigmoid', dropout=0, fusion_type='cat')
    torch.manual_seed(0)
    model = P.DPRNNSpeTasNet(**kw).eval().cuda()
    model.precision = 'bf16'
    model.n_streams = a.streams
    mix = (0.05 * torch.randn(a.batch, 24000)).cuda()
    ref = (0.05 * torch.randn(a.batch, 24000)).cuda()
    rl = torch.tensor(24000.)
    L = P.lib()
    with torch.no_grad():
        for _ in range(3):
            model(mix, ref, rl)
        torch.cuda.synchronize()
        L.timing = {n: [] for n in L.protos}
        L.timeline = []
        base = torch.cuda.Event(enable_timing=True)
        base.record()
        model(mix, ref, rl)
        torch.cuda.synchronize()
    rows = []
    for name, evs in L.timing.items():
        for (e0, e1) in evs:
            rows.append((base.elapsed_time(e0), base.elapsed_time(e1), name.replace('dprnn_', '')))
    rows.sort()
    end = max(r[1] for r in rows)
    print(f'step {end:.2f} ms, {len(rows)} launches')
    busy = {}
    for s, e, n in rows:
        busy[n] = busy.get(n, 0.0) + (e - s)
    for n, v in sorted(busy.items(), key=lambda kv: -kv[1])[:8]:
        print(f'  {n:28s} sum of launch spans {v:8.2f} ms')
    lstm = [(s, e) for s, e, n in rows if n.startswith('lstm')]
    # time during which at least one LSTM kernel is in flight, and during which none is
    ev = sorted([(s, 1) for s, e in lstm] + [(e, -1) for s, e in lstm])
    depth, last, cover, multi = 0, 0.0, 0.0, 0.0
    for t, d in ev:
        if depth >= 1:
            cover += t - last
        if depth >= 2:
            multi += t - last
        depth += d
        last = t
    print(f'  >=1 LSTM kernel in flight {cover:.2f} ms, >=2 in flight {multi:.2f} ms, none {end - cover:.2f} ms')
    for s, e, n in rows[:a.rows]:
        print(f'{s:8.3f} -> {e:8.3f}  ({e - s:6.3f})  {n}')


if __name__ == '__main__':
    main()
