#!/bin/bash
# LSTM scheduling variants of the headline bench (one GPU):  gpurun -- 'bash tools/sweep_lstm.sh'
O=gpurun_out; mkdir -p $O
Q="--no-cpu-baseline --modes 0 --gpu-reference 0 --cfg5 0 --steps 5 --warmup 3"
for v in "3 1 0" "1 1 0" "1 0 0" "3 0 0" "2 0 0" "2 0 37" "1 3 0" "1 4 0"; do
  set -- $v
  timeout 200 python bench.py $Q --streams $1 --lstm-slices $2 --lstm-pairs $3 > $O/sw_$1_$2_$3.json 2> $O/sw_$1_$2_$3.err
  python - "$O/sw_$1_$2_$3.json" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d['kernels']
    l = {n: round(v['ms_avg'], 3) for n, v in k.items() if 'lstm' in n}
    print(f"streams/slices/pairs {sys.argv[2]}: {d['ms_per_step']:.2f} ms/step  {d['value']:.0f} audio-s/s  roofline {d['roofline']['frac']:.3f}  lstm {l}")
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
