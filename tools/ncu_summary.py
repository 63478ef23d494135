"""Summarise an `ncu --set full` report into one row per kernel (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_ncu_kernels.md

Per kernel name: launches, mean duration, DRAM bytes read+written per launch, achieved DRAM GB/s
(traffic / duration) and its fraction of MEASURED_PEAKS.json hbm_gbs, ncu's own DRAM-throughput %,
tensor-pipe active %, XU (MUFU) pipe %, registers/thread, dynamic shared memory.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    'gpu__time_duration.sum': 'dur',
    'dram__bytes_read.sum': 'rd',
    'dram__bytes_write.sum': 'wr',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'xu_pct',
    'dram__bytes.sum.per_second': 'bps',
    'launch__registers_per_thread': 'regs',
    'launch__shared_mem_per_block_dynamic': 'smem',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ_pct',
}
UNIT = {'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 'second': 1.0, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0,
        'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
        'byte/s': 1.0, 'Kbyte/s': 1e3, 'Mbyte/s': 1e6, 'Gbyte/s': 1e9, 'Tbyte/s': 1e12,
        'byte/block': 1.0, 'Kbyte/block': 1e3, 'Mbyte/block': 1e6}


def main():
    rep = sys.argv[1]
    if rep.endswith('.csv'):      # already exported with `ncu -i X.ncu-rep --page raw --csv`
        raw = open(rep).read()
    else:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kname = col['Kernel Name']
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    hbm = peaks.get('hbm_gbs', 6650.0)
    agg = OrderedDict()
    for r in data:
        name = re.sub(r'\(.*', '', r[kname]).replace('void ', '').replace('dprnn::', '').strip()
        d = agg.setdefault(name, {'n': 0})
        d['n'] += 1
        for m, key in WANT.items():
            if m not in col:
                continue
            try:
                v = float(r[col[m]].replace(',', ''))
            except ValueError:
                continue
            v *= UNIT.get(units[col[m]], 1.0) if key in ('dur', 'rd', 'wr', 'smem', 'bps') else 1.0
            d[key] = d.get(key, 0.0) + v
    print(f'| kernel | launches | avg ms | DRAM rd+wr / launch (GB) | DRAM GB/s | of measured {hbm:.0f} GB/s | ncu dram % | '
          'tensor pipe % | XU pipe % | SM thr % | regs | dyn smem KB | grid x block |')
    print('|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|')
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1].get('dur', 0)):
        n = d['n']
        g = lambda k: d.get(k, 0.0) / n
        dur = g('dur')
        traffic = g('rd') + g('wr')
        if traffic == 0.0:            # light section sets carry the rate, not the byte counters
            traffic = g('bps') * dur
        gbs = traffic / dur / 1e9 if dur else 0.0
        print(f"| `{name}` | {n} | {dur * 1e3:.4f} | {traffic / 1e9:.4f} | {gbs:.0f} | {gbs / hbm:.2f} | {g('dram_pct'):.1f} | "
              f"{g('tensor_pct'):.1f} | {g('xu_pct'):.1f} | {g('sm_pct'):.1f} | {g('regs'):.0f} | {g('smem') / 1e3:.1f} | "
              f"{g('grid'):.0f} x {g('block'):.0f} |")


if __name__ == '__main__':
    main()
