#!/bin/bash
# The bench lines committed under profiles/ (one GPU):  gpurun --timeout 1200 -- 'bash tools/run_final_benches.sh'
set -u
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; timeout 400 python bench.py "$@" > $O/final_$name.json 2> $O/final_$name.err || echo "FAILED $name"; tail -c 300 $O/final_$name.err; }
run cfg2_bf16 --steps 5 --warmup 3
run cfg2_fp32 --steps 2 --warmup 3 --precision fp32 --no-cpu-baseline
run cfg1_bf16 --workload cfg1 --steps 10 --warmup 3
run cfg3_bf16 --workload cfg3 --steps 5 --warmup 3
run cfg4_bf16 --workload cfg4 --steps 5 --warmup 3
run cfg5_tc --workload cfg5 --steps 3 --warmup 3
run cfg5_fp32 --workload cfg5 --steps 2 --warmup 3 --precision fp32 --no-cpu-baseline
run cfg2_reference --impl reference --steps 2 --warmup 1
for f in $O/final_*.json; do echo "== $f"; python tools/bench_summary.py < $f 2>/dev/null | head -3; done
