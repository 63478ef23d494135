#!/bin/bash
# GPU suite + default bench at HEAD, then the A/B of Engine.fold_fused:  gpurun --timeout 300 -- 'bash tools/verify_fold_fused.sh'
set -u
O=gpurun_out
mkdir -p $O
t0=$(date +%s)
timeout 200 python -m pytest tests -m gpu -x -q > $O/ff_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/ff_tests.log)  [$(( $(date +%s) - t0 )) s]"
grep -E "FAILED|Error|assert" $O/ff_tests.log | head -20
Q="--steps 20 --warmup 5 --no-cpu-baseline --modes 0 --gpu-reference 0 --cfg5 0"
timeout 100 python bench.py $Q > $O/ff_bench_fused.json 2> $O/ff_bench_fused.err; echo "fused rc=$?"; python tools/bench_summary.py 8 < $O/ff_bench_fused.json
DPRNN_FOLD_FUSED=0 timeout 100 python bench.py $Q > $O/ff_bench_unfused.json 2> $O/ff_bench_unfused.err; echo "unfused rc=$?"; python tools/bench_summary.py 8 < $O/ff_bench_unfused.json
timeout 100 python bench.py $Q > $O/ff_bench_fused2.json 2> $O/ff_bench_fused2.err; echo "fused again rc=$?"; python tools/bench_summary.py 3 < $O/ff_bench_fused2.json
echo "[$(( $(date +%s) - t0 )) s]"
