"""Condense an `ncu --set full` raw-page CSV (ncu -i X.ncu-rep --page raw --csv) to the counters the design is argued
from, one block per profiled launch (the format bench.py's ncu_traffic_per_launch reads).

    python tools/ncu_pick.py gpurun_out/p_train_top_raw.csv > profiles/rN_train_top_ncu_full.txt
"""
import csv
import io
import sys

PICK = ['launch__grid_size', 'launch__block_size', 'launch__cluster_dim_x', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__waves_per_multiprocessor']


def main():
    rows = list(csv.reader(io.StringIO(open(sys.argv[1]).read())))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print(f"Kernel Name = {r[col['Kernel Name']]} ")
        for m in PICK:
            if m in col:
                print(f'{m} = {r[col[m]]} {units[col[m]]}'.rstrip())
        print()


if __name__ == '__main__':
    main()
