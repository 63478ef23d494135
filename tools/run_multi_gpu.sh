#!/bin/bash
# Multi-GPU bench lines (one node):  gpurun --gpus N --timeout 1500 -- 'bash tools/run_multi_gpu.sh N r2'
set -u
N=${1:-2}; R=${2:-r2}; WHICH=${3:-all}          # WHICH: all | train (cfg 5 only)
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; port=$2; shift 2; timeout 600 $TR --master-port $port bench.py --gpus $N "$@" > $O/${R}_bench_${name}_${N}gpu.json 2> $O/${R}_bench_${name}_${N}gpu.err || echo "FAILED $name"; grep -h '^{' $O/${R}_bench_${name}_${N}gpu.json | tail -1 > $O/tmp.json; mv $O/tmp.json $O/${R}_bench_${name}_${N}gpu.json; }
if [ "$WHICH" = all ]; then
run cfg3_full 29701 --workload cfg3 --full --no-cpu-baseline
run cfg4 29702 --workload cfg4 --steps 5 --no-cpu-baseline
fi
run cfg5 29703 --workload cfg5 --steps 3 --no-cpu-baseline
if [ "$WHICH" = all ]; then
run cfg2 29704 --steps 5 --no-cpu-baseline
fi
for f in $O/${R}_bench_*_${N}gpu.json; do echo "== $f"; python - $f <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read())
print(f"{d['n_gpus']} GPUs: {d['ms_per_step']:.2f} ms/step  value {d['value']:.1f}  e2e {d['e2e']['value']:.1f}  ranks {d['ranks']}")
for k in ('cfg5', 'train'):
    if d.get(k): print('  ', k, d[k])
PY
done
