#!/bin/bash
# The bench lines committed under profiles/ (one GPU):  gpurun --timeout 1500 -- 'bash tools/run_final_benches.sh r2'
set -u
R=${1:-r2}
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; timeout 500 python bench.py "$@" > $O/${R}_bench_$name.json 2> $O/${R}_bench_$name.err || echo "FAILED $name"; tail -c 300 $O/${R}_bench_$name.err; }
run cfg2_fp16 --steps 8 --warmup 3
run cfg2_fp16_1stream_sliced --steps 8 --warmup 3 --streams 1 --lstm-slices 0 --no-cpu-baseline --modes 0 --gpu-reference 0 --cfg5 0
run cfg2_fp16_fp32master --steps 8 --warmup 3 --residual-bf16 0 --no-cpu-baseline --modes 0 --gpu-reference 0 --cfg5 0
run cfg2_bf16 --steps 8 --warmup 3 --precision bf16 --no-cpu-baseline --modes 0 --gpu-reference 0 --cfg5 0
run cfg2_fp32 --steps 2 --warmup 3 --precision fp32 --no-cpu-baseline --modes 0 --gpu-reference 0 --cfg5 0
run cfg1_fp16 --workload cfg1 --steps 10 --warmup 3
run cfg3_fp16 --workload cfg3 --steps 5 --warmup 3
run cfg4_fp16 --workload cfg4 --steps 5 --warmup 3
run cfg5_tc --workload cfg5 --steps 3 --warmup 3
run cfg2_reference --impl reference --steps 2 --warmup 1
for f in $O/${R}_bench_*.json; do echo "== $f"; python tools/bench_summary.py < $f 2>/dev/null | head -3; done
