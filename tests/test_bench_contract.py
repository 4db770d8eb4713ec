"""
The bench contract on the CPU: the keys of the committed bench line (profiles/r02_bench_n1.json, written by bench.py
on a B200) and a live run of the reference arm (the oracle port on the host cores), which needs no GPU.
"""
import json, os, subprocess, sys

from conftest import ROOT

BASE_KEYS = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
             'dtype', 'data', 'config')


def test_committed_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, 'profiles', 'r02_bench_n1.json')))
    for k in BASE_KEYS + ('e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline', 'sustained', 'secondary', 'batch96'):
        assert k in d, k
    assert d['n_gpus'] == 1 and d['warmup'] >= 3 and d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'configs[1]' in d['config']['workload'] and 'model' not in d['config']
    e = d['e2e']
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step', 'mode', 'stats', 'full', 'replay'} <= set(e)
    assert e['mode'] == 'stats' and e['value'] == e['stats']['value'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert e['value'] < d['value'] and e['value'] >= 0.85 * d['value']              # end to end tracks the resident pipeline
    assert e['full']['d2h_bytes_per_step'] > 32 * d['draws_per_step_per_gpu'] and e['replay']['h2d_bytes_per_step'] == 16 * d['draws_per_step_per_gpu']
    assert d['gpu_launches'] >= 3 * d['steps']                                      # sampler, draw kernel, reductions in every step
    r = d['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'kernel_share_of_step'} <= set(r) and r['bound'] == 'hbm' and r['unit'] == 'GB/s'
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and 0.9 < r['frac'] < 1.2 and 0.7 < r['kernel_share_of_step'] < 1
    c = d['cpu_baseline']
    assert {'value', 'unit', 'cores', 'kind', 'sample', 'port'} <= set(c) and c['kind'] == 'reference' and c['cores'] == 1
    assert 'popcount_lk' in c['popcount'] and c['port']['value'] > c['value']        # the port (int.bit_count) is the faster CPU arm
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(d['clocks'])
    assert not {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'} & set(d['clocks']['reasons'])
    assert d['parity_max_rel_err_vs_oracle'] <= 1e-12
    s = d['sustained']
    assert s['steps'] >= 500 and s['p10_ms'] <= s['median_ms_per_step'] <= s['p90_ms']
    legs = {leg['workload'].split(':')[0] for leg in d['secondary']}
    assert {'configs[2]', 'configs[3]', 'configs[4] shape'} <= legs
    for leg in d['secondary']:
        # int_frac is word AND+POPC per second over the measured DIRECT peak; the carry-save counting of k_small / k_small_win
        # trades POPCs for LOP3s, so more than 1 is expected (k_small_win: 1.6) -- twice the peak would be a units error
        assert leg['value'] > 0 and leg['ms_per_step'] > 0 and 0 < leg['int_frac'] < 2.0
    assert d['batch96']['window_likelihoods'] == 96 * d['config']['windows_per_gpu'] * 1000 and d['batch96']['statistics_finite']
    for n in (2, 4, 8):                                                              # the scaling lines of the same round
        m = json.load(open(os.path.join(ROOT, 'profiles', 'r02_bench_n%d.json' % n)))
        assert m['n_gpus'] == n and m['value'] > 0.9 * n * d['value'] and m['e2e']['value'] > 0.8 * n * e['value']
        assert m['strong']['scaling'] == 'strong' and m['strong']['allreduce_ms_median'] > 0 and m['strong']['statistics_finite']


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
    ours = json.load(open(os.path.join(ROOT, 'profiles', 'r02_bench_n1.json')))
    assert d['metric'] == ours['metric'] and d['unit'] == ours['unit'] and d['config'] == ours['config']
