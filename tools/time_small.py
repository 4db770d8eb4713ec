#!/usr/bin/env python3
""" Times the draw kernel alone on the bench workload (configs[1] embryo):
python tools/time_small.py [n] [model] [steps] [windows] [subsamples] [depth].
Kernel variants are chosen by the environment (e.g. LDPGTA_SMALL_TMA=1). """
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import ldpgta_b200 as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = sys.argv[2] if len(sys.argv) > 2 else 'homogeneous'
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
windows = int(sys.argv[4]) if len(sys.argv) > 4 else 0
subs = int(sys.argv[5]) if len(sys.argv) > 5 else 32
depth = float(sys.argv[6]) if len(sys.argv) > 6 else 0.1
dev = torch.device('cuda', 0)
wins, reads = bench.workload_shape(depth, 12345, min_reads=n + 2, max_windows=windows)
sizes, kind, prop = ([bench.N_HAP], L.HOMOGENEOUS, None) if model == 'homogeneous' else \
    ([2504, 2504], L.RECENT_ADMIXTURE, None) if model == 'recent' else ([2504, 2504], L.DISTANT_ADMIXTURE, [0.6, 0.4])
eng = L.Engine(sizes)
keep = []
for g, nh in enumerate(sizes):
    m = bench.synth_masks(torch, dev, reads, nh, 12345 + 77 * g)
    eng.attach_read_masks_dev(g, m.data_ptr(), int(reads.sum()), keepalive=m)
idx, _ = bench.plan_draws(reads, subs, n, 12346)
d_idx = torch.from_numpy(idx).to(dev)
d_out = torch.empty((len(idx), 4), dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(eng.stream, device=dev)
def step():
    eng.likelihoods_uniform_dev(kind, n, d_idx.data_ptr(), len(idx), d_out.data_ptr(), proportions=prop)
for _ in range(min(20, steps)):
    step()
eng.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record(stream)
for k in range(steps):
    step(); ev[k + 1].record(stream)
eng.synchronize()
ms = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(steps)])
print(json.dumps({'n': n, 'model': model, 'env': {k: v for k, v in os.environ.items() if k.startswith('LDPGTA_')}, 'median_ms': float(np.median(ms)),
                  'p10': float(np.percentile(ms, 10)), 'p90': float(np.percentile(ms, 90)), 'draws': len(idx), 'draws_per_s': len(idx) / (float(np.median(ms)) * 1e-3), 'wordop_frac': ((1 << n) - 1) * sum((h + 63) // 64 for h in sizes) * len(idx) / (float(np.median(ms)) * 1e-3) / 2.2797e12, 'checksum': float(d_out.sum().item())}))
