#!/usr/bin/env python3
"""Times the phases of the drop-in bootstrap() on BASELINE.json configs[0]: chr21-sized synthetic disomy
sample at 0.05x, 1 000 000-SNP x 5008-haplotype random panel, on one GPU.
The reference's own phase times for the same input (measured in the survey container, 1 core):
supporting_dictionaries 2.39 s, build_hap_dict 0.75 s, build_gw_dict 0.05 s, bootstrap loop 1.19 s,
statistics 0.26 s."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ldpgta_b200 as L
from ldpgta_b200 import driver

L_CHR21, S, N = 46709983, int(os.environ.get('SNPS', 1000000)), 5008
random.seed(12345)
t0 = time.time()
positions = sorted(random.sample(range(1, L_CHR21 + 1), S))
refalt = [tuple(random.sample('ACGT', 2)) for _ in range(S)]
rows = [random.getrandbits(N) for _ in range(S)]
leg = tuple(('chr21', p, r, a) for p, (r, a) in zip(positions, refalt))
hap = {'ALL': tuple(rows)}
nh = {'ALL': N}
hapl = [{p: (ra[1] if (row >> b) & 1 else ra[0]) for p, row, ra in zip(positions, rows, refalt)} for b in (17, 4242)]
# the simulator's read model (MIX_HAPLOIDS.build_obs_tab: uniform starts, 36 bp, disomy), restated
random.seed(777)
n_reads = int(0.05 * L_CHR21 // 36)
obs = []
for i in range(n_reads):
    p = random.randrange(L_CHR21) + 1
    h = random.choices((0, 1), k=1)[0]
    rid = '%d.%d.%s.%d' % (p - 18, p + 17, 'AB'[h], i)
    src = hapl[h]
    obs.extend((q, rid, src[q]) for q in range(p - 18, p + 18) if q in src)
obs.sort(key=lambda r: r[0])
t_gen = time.time() - t0
print('generated %d observations in %.1f s' % (len(obs), t_gen), file=sys.stderr)

phases = {}
def timed(name, fn, *a, **k):
    t = time.time(); r = fn(*a, **k); phases[name] = phases.get(name, 0.0) + time.time() - t; return r

random.seed(1)
warm = timed('cuda_context_first_use (once per process)', L.Engine, [64]); warm.close()
examine = timed('model_init (build_hap_dict + upload)', driver.make_examiner, obs, leg, hap, nh, {'ALL'}, L.ImplicitModels())
sd = timed('supporting_dictionaries (scores on device)', driver.supporting_dictionaries, obs, leg, hap, nh, {'ALL'}, 0.05, examine=examine)
gw = timed('build_gw_dict', driver.build_gw_dict, sd, 100000, 0, 6, 4, 2)
draws = timed('plan_draws (random.sample)', driver.plan_draws, gw, 6, 4, 32)
lik = timed('evaluate_draws (masks + batch on device)', driver.evaluate_draws, examine, sd.reads, draws)
st = timed('statistics (device)', driver.statistics, lik, gw)
examine.close()
n_eval = sum(len(v) for v in lik.values())

# the array-native route (ldpgta_b200/packed.py): the panel is uploaded once per chromosome and shared by samples
import numpy as np
from ldpgta_b200.packed import PackedSample
packed = {}
def timed2(name, fn, *a, **k):
    t = time.time(); r = fn(*a, **k); packed[name] = packed.get(name, 0.0) + time.time() - t; return r
panel = timed2('resident_panel (once per chromosome, shared by all samples)', L.ResidentPanel, leg, hap, nh)
timed2('packed_legend (once per chromosome)', lambda: panel.legend)
per_sample = []
for rep in range(3):
    t = time.time()
    phases2 = {}
    t1 = time.time(); ps = PackedSample(obs, panel, {'ALL'}, 0.05); phases2['pack + gather + read masks + scores'] = time.time() - t1
    t1 = time.time(); res = ps.run(100000, 32, 0, 6, 4, 2, rng=np.random.default_rng(rep)); phases2['windows + draws + likelihoods + statistics'] = time.time() - t1
    t1 = time.time(); ref_lik, ref_gw, _ = res.as_reference(); phases2['unfold into the reference dicts (optional)'] = time.time() - t1
    ps.close()
    per_sample.append({'total_s': time.time() - t, 'phases_s': phases2})
# config 5's regime: 1000 subsamples per window, draws made on the device (only the window lists go up)
deep = []
for rep in range(3):
    t = time.time()
    ps = PackedSample(obs, panel, {'ALL'}, 0.05)
    t1 = time.time()
    res1000 = ps.run(100000, 1000, 0, 6, 4, 2, seed=rep, want_draws=False, keep_pinned=True)
    t2 = time.time()
    ps.close()
    deep.append({'total_s': time.time() - t, 'pack_s': t1 - t, 'run_s': t2 - t1, 'window_likelihoods': int(res1000.draw_offsets[-1])})
    res1000.release()
packed['subsamples_1000_device_draws'] = deep
# the same from observations packed once (PackedObservations.save / load): no walk over the pickled tuples
from ldpgta_b200.packed import pack_observations, PackedObservations
pack_observations(obs, panel.legend).save('/tmp/chr21_sample.npz')
arr = []
for rep in range(3):
    t = time.time()
    po = PackedObservations.load('/tmp/chr21_sample.npz')
    t0 = time.time()
    ps = PackedSample(po, panel, {'ALL'}, 0.05)
    t1 = time.time()
    r = ps.run(100000, 1000, 0, 6, 4, 2, seed=rep, want_draws=False, keep_pinned=True)
    t2 = time.time()
    ps.close(); r.release()
    arr.append({'total_s': time.time() - t, 'load_npz_s': t0 - t, 'gather_masks_scores_s': t1 - t0, 'run_s': t2 - t1})
packed['subsamples_1000_from_npz'] = arr
assert list(ref_gw) == list(gw) and list(ref_lik) == list(lik)
assert all(set(a) == set(b) for a, b in zip(ref_gw.values(), gw.values()))
packed['per_sample_runs'] = per_sample
packed['BPH_vs_SPH'] = res.llr_per_chromosome()[('BPH', 'SPH')]

print(json.dumps({'windows': len(gw), 'evaluated_windows': len(lik), 'window_likelihoods': n_eval, 'phases_s': phases,
                  'total_s': sum(phases.values()),
                  'BPH_vs_SPH': st['LLRs_per_chromosome'][('BPH', 'SPH')], 'packed_route': packed}, indent=1))
