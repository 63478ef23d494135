#!/bin/bash
# A/B of the training step (cfg 5) knobs:  gpurun --timeout 900 -- 'bash tools/ab_cfg5.sh'
set -u
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $O/ab_cfg5_$name.json 2> $O/ab_cfg5_$name.err || echo "FAILED $name"; 
  python - <<PY
import json
try:
    d = json.load(open('$O/ab_cfg5_$name.json'))
    pk = d.get('kernels', {})
    top = sorted(pk.items(), key=lambda kv: -kv[1]['ms_total'])[:14]
    print('$name', round(d['ms_per_step'], 2), 'ms/step', [(k.replace('dprnn_', ''), round(v['ms_avg'], 3), v['launches']) for k, v in top])
except Exception as e:
    print('$name', 'no line', e)
PY
}
run default X=1
