"""The bench lines committed under profiles/ carry every key of the driver's contract (bench.py prints one JSON line)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r1_bench_*.json')))


@pytest.mark.parametrize('path', LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_bench_line_has_the_contract_keys(path):
    line = json.loads(open(path).read().strip().splitlines()[-1])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches'):
        assert k in line, k
    assert line['unit'] == 'audio-s/s' and line['higher_is_better'] is True and line['vs_baseline'] is None
    assert isinstance(line['config'].get('workload'), str)
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(line['e2e'])
    if line.get('impl') == 'reference':
        assert line['gpu_launches'] == 0 and line['cpu_baseline']['kind'] in ('reference', 'port')
        return
    assert line['gpu_launches'] > 0 and line['value'] > 0
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(line['clocks'])
    for bad in ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'):
        assert bad not in line['clocks']['reasons']
    roof = line['roofline']
    assert roof['bound'] in ('hbm', 'tensor') and roof['unit'] in ('GB/s', 'TFLOP/s')
    assert abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-6 and 'traffic' in roof
    assert line['steps'] >= 1 and line['warmup'] >= 3
