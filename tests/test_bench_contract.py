"""
The bench contract on the CPU: the keys of the committed bench line (profiles/r01_bench_n1.json, written by bench.py
on a B200) and a live run of the reference arm (the oracle port on the host cores), which needs no GPU.
"""
import json, os, subprocess, sys

from conftest import ROOT

BASE_KEYS = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
             'dtype', 'data', 'config')


def test_committed_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, 'profiles', 'r01_bench_n1.json')))
    for k in BASE_KEYS + ('e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['n_gpus'] == 1 and d['warmup'] >= 3 and d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'configs[1]' in d['config']['workload'] and 'model' not in d['config']
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(d['e2e']) and d['e2e']['h2d_bytes_per_step'] > 0
    assert d['e2e']['value'] < d['value'] and d['gpu_launches'] >= d['steps']
    r = d['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(r) and r['bound'] == 'hbm' and r['unit'] == 'GB/s'
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and 0 < r['frac'] < 1.2
    c = d['cpu_baseline']
    assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(c) and c['kind'] == 'port' and c['cores'] == 1
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(d['clocks'])
    assert not {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'} & set(d['clocks']['reasons'])
    assert d['parity_max_rel_err_vs_oracle'] <= 1e-12


def test_reference_arm_runs_without_a_gpu():
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                    # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['value'] > 0
    for k in BASE_KEYS + ('cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    import make_ref                                            # oracle/make_ref.py: the staged, unmodified reference scripts
    assert d['cpu_baseline']['kind'] == ('reference' if make_ref.available() else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    ours = json.load(open(os.path.join(ROOT, 'profiles', 'r01_bench_n1.json')))
    assert d['metric'] == ours['metric'] and d['unit'] == ours['unit'] and d['config'] == ours['config']
