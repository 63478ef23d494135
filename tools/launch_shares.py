"""Per-kernel share of the GPU time from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_shares.py gpurun_out/launches.csv > profiles/rN_launch_shares.md
"""
import csv
import io
import re
import sys
from collections import OrderedDict


def main():
    raw = open(sys.argv[1]).read()
    raw = raw[raw.index('"ID"'):]
    agg = OrderedDict()
    for r in csv.DictReader(io.StringIO(raw)):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('dprnn::', '').strip()
        if name.startswith('at::') or name.startswith('cub::') or 'elementwise' in name:
            name = 'torch: ' + name.split('<')[0]
        v = float(r['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(r['Metric Unit'], 1e-6)
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v
    total = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f'| kernel | launches | total ms | share |\n|---|---:|---:|---:|')
    for name, (k, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{name}` | {k} | {ms:.3f} | {100 * ms / total:.1f}% |')
    print(f'| **all** | {n} | {total:.3f} | 100% |')


if __name__ == '__main__':
    main()
