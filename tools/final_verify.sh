#!/bin/bash
# HEAD verification in one short GPU session (one GPU, ~5 min):  gpurun --timeout 420 -- 'bash tools/final_verify.sh'
#   1. the whole GPU parity suite, 2. the default bench line (what the driver runs), 3. the cfg-5 training-step line,
#   4. a full ncu capture of the two recurrent training kernels (after the same command exited 0 without ncu).
set -u
O=gpurun_out
mkdir -p $O
t0=$(date +%s)
timeout 260 python -m pytest tests -m gpu -x -q > $O/fv_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/fv_tests.log)  [$(( $(date +%s) - t0 )) s]"
timeout 240 python bench.py > $O/fv_bench_default.json 2> $O/fv_bench_default.err; echo "bench default rc=$?  [$(( $(date +%s) - t0 )) s]"
python tools/bench_summary.py 6 < $O/fv_bench_default.json
timeout 120 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $O/fv_bench_cfg5.json 2> $O/fv_bench_cfg5.err; echo "bench cfg5 rc=$?  [$(( $(date +%s) - t0 )) s]"
python tools/bench_summary.py 14 < $O/fv_bench_cfg5.json
timeout 90 python tools/profile_train_step.py > $O/fv_plain_train.log 2>&1 && \
timeout 150 ncu --set full --clock-control none --import-source on -k "regex:^(lstm_bptt|lstm_tc)" -c 4 \
    -o /tmp/fv_train_top -f python tools/profile_train_step.py > $O/fv_ncu_train_top.log 2>&1
ncu -i /tmp/fv_train_top.ncu-rep --page raw --csv > $O/fv_train_top_raw.csv 2>/dev/null
ncu -i /tmp/fv_train_top.ncu-rep --page details > $O/fv_train_top_details.txt 2>/dev/null
echo "ncu done [$(( $(date +%s) - t0 )) s]"; ls -la $O/fv_* | tail -12
