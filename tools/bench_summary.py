"""Pretty-print the JSON line of bench.py (stdin): headline numbers + the top kernels by time."""
import json
import sys

line = [l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]
d = json.loads(line)
r = d.get('roofline') or {}
print(f"ms/step {d['ms_per_step']:.2f}  value {d['value']:.1f}  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}  "
      f"roofline.frac {r.get('frac', 0):.3f}  cfg {d['config'].get('streams')} streams")
top = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for k, v in sorted(d.get('kernels', {}).items(), key=lambda kv: -kv[1]['ms_total'])[:top]:
    print(f"   {k:32s} n={v['launches']:4d} total {v['ms_total']:8.3f} ms  avg {v['ms_avg']:.4f}")
