#!/usr/bin/env python3
""" One model.get_likelihoods(alleles) call -- the operator ANEUPLOIDY_TEST.bootstrap invokes per draw (:285) -- timed for
4 / 8 / 12 / 16 reads on a 5008-haplotype panel: the GPU class of this package beside the reference's own class (oracle/_ref,
unmodified, on one host core).  python tools/operator_latency.py > gpurun_out/operator_latency.json """
import contextlib, io, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import bench
import ldpgta_b200 as L

rng = random.Random(5)
N, S = 5008, 4000
founders = [rng.getrandbits(N) for _ in range(8)]
rows = [founders[rng.randrange(8)] ^ (rng.getrandbits(N) & rng.getrandbits(N) & rng.getrandbits(N)) for _ in range(S)]     # LD-like: 8 founders + 12 % noise
leg = [('chr1', 100 + 37 * i, 'A', 'C') for i in range(S)]
obs = [(100 + 37 * i, 'r%d' % (i // 2), 'C' if rng.random() < 0.5 else 'A') for i in range(S)]
reads = {}
for pos, rid, base in obs:
    reads.setdefault(rid, []).append((pos, base))
ids = list(reads)
with contextlib.redirect_stdout(io.StringIO()):
    gpu = L.homogeneous(obs, leg, {'ALL': rows}, {'ALL': N}, L.ImplicitModels())
    cls, kind, pc = bench.cpu_model_class()
    import ldpgta_oracle as O
    cpu = cls(obs, leg, {'ALL': rows}, {'ALL': N}, O.make_models(8) if kind == 'reference' else {})
out = {'haplotypes': N, 'cpu_kind': kind, 'cpu_popcount': pc, 'sizes': {}}
for n in (4, 8, 12, 16):
    draws = [tuple(reads[r] for r in rng.sample(ids[k * 40:(k + 1) * 40], n)) for k in range(40)]
    for d in draws[:5]:
        gpu.get_likelihoods(d)
    t0 = time.perf_counter()
    got = [gpu.get_likelihoods(d) for d in draws]
    t_gpu = (time.perf_counter() - t0) / len(draws)
    rec = {'gpu_us_per_call': 1e6 * t_gpu}
    if n <= 8:                                           # the reference needs MODELS tables beyond 4 reads; BUILD(8) is quick
        k = len(draws) if n <= 4 else 10
        t0 = time.perf_counter()
        ref = [cpu.get_likelihoods(d) for d in draws[:k]]
        rec['cpu_us_per_call'] = 1e6 * (time.perf_counter() - t0) / k
        rec['max_rel_diff'] = max(abs(a - b) / max(abs(b), 1e-300) for g, r in zip(got, ref) for a, b in zip(g, r))
    out['sizes'][n] = rec
gpu.close()
print(json.dumps(out, indent=1))
