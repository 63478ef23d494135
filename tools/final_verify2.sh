#!/bin/bash
# Last GPU session of the round (one GPU, ~3.5 min):  gpurun --timeout 260 -- 'bash tools/final_verify2.sh'
# GPU suite, the default bench line, then the ragged (cfg 3) A/B of Engine.fold_fused and the cfg 1 / cfg 4 lines at HEAD.
set -u
O=gpurun_out
mkdir -p $O
t0=$(date +%s)
timeout 200 python -m pytest tests -m gpu -x -q > $O/fv2_tests.log 2>&1; echo "tests rc=$? $(tail -1 $O/fv2_tests.log)  [$(( $(date +%s) - t0 )) s]"
grep -E "FAILED|Error|assert" $O/fv2_tests.log | head -20
timeout 150 python bench.py > $O/fv2_bench_default.json 2> $O/fv2_bench_default.err; echo "default rc=$?  [$(( $(date +%s) - t0 )) s]"; python tools/bench_summary.py 12 < $O/fv2_bench_default.json
Q="--no-cpu-baseline"
timeout 60 python bench.py --workload cfg3 --steps 5 --warmup 3 $Q > $O/fv2_bench_cfg3.json 2> $O/fv2_bench_cfg3.err; echo "cfg3 rc=$?  [$(( $(date +%s) - t0 )) s]"; python tools/bench_summary.py 6 < $O/fv2_bench_cfg3.json
DPRNN_FOLD_FUSED=0 timeout 60 python bench.py --workload cfg3 --steps 5 --warmup 3 $Q > $O/fv2_bench_cfg3_unfused.json 2> $O/fv2_bench_cfg3_unfused.err; echo "cfg3 unfused rc=$?"; python tools/bench_summary.py 3 < $O/fv2_bench_cfg3_unfused.json
timeout 40 python bench.py --workload cfg1 --steps 10 --warmup 3 $Q > $O/fv2_bench_cfg1.json 2> $O/fv2_bench_cfg1.err; echo "cfg1 rc=$?"; python tools/bench_summary.py 3 < $O/fv2_bench_cfg1.json
timeout 40 python bench.py --workload cfg4 --steps 5 --warmup 3 $Q > $O/fv2_bench_cfg4.json 2> $O/fv2_bench_cfg4.err; echo "cfg4 rc=$?"; python tools/bench_summary.py 3 < $O/fv2_bench_cfg4.json
echo "[$(( $(date +%s) - t0 )) s]"
