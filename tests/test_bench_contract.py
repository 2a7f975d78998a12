"""bench.py's output contract: the committed end-of-round line (profiles/r1_bench_final.json) carries every key the
driver reads, and the reference arm (`--impl reference`, the oracle port on the host cores) runs and prints its line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e'}


def test_committed_bench_line_has_the_contract_keys():
    line = json.load(open(os.path.join(ROOT, 'profiles', 'r1_bench_final.json')))
    assert BASE_KEYS | {'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'} <= set(line)
    assert 'workload' in line['config'] and 'model' not in line['config']
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(line['e2e'])
    assert line['e2e']['h2d_bytes_per_step'] > 0 and line['e2e']['d2h_bytes_per_step'] > 0
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(line['roofline'])
    assert abs(line['roofline']['frac'] - line['roofline']['achieved'] / line['roofline']['peak']) < 1e-3
    assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(line['cpu_baseline'])
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(line['clocks'])
    assert line['gpu_launches'] > 0 and line['vs_baseline'] is None and line['higher_is_better'] is True
    assert abs(line['value'] - 32 * line['n_gpus'] / (line['ms_per_step'] * 1e-3)) / line['value'] < 1e-3


def test_reference_arm_runs_on_the_host_and_prints_its_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert BASE_KEYS | {'impl', 'cpu_baseline'} <= set(line)
    assert line['impl'] == 'reference' and line['steps'] == 1 and line['value'] > 0
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
