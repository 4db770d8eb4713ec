#!/usr/bin/env python3
"""
tools/sweep_sizes.py [model] -- every draw size 2..18 of one model class on one GPU: the draw kernel alone (CUDA events,
median of the steps), draws/s, and three rooflines:

  popc   algorithmic AND+POPC word-ops (G * (2^n - 1) * W per draw) against the measured direct AND+POPC peak
         (profiles/INT_PEAKS.json: wordop64_and_popc_add)
  two    the two-term ceiling for 5 or more reads: time of the popcounts when three words share two POPC.64 through a
         carry-save adder (the counting loops of k_lanes / k_general: 4/3 POPC.32 per word-op against the measured POPC.32
         peak) PLUS the time of the partition polynomial's multiply-adds against the measured IMAD.32 / DFMA peaks -- the
         two phases run one after the other inside a CTA.  Multiply-adds per draw: (3^(n-5) - 1) / 2 colourings x
         (81 + 2 * 16 / R) for one population (conv16, and a 64-bit dot16 per run of R colourings that share the X row:
         R = 8 from 15 reads, 4 at 12-14, 1 below), x (4 * 81 + 4 * 32 / R) for recent admixture, x (4 * 81 + 6 * 32 / R)
         for distant admixture of two groups (the same integer convolutions, six dot products); see csrc/poly_general.cu.
  hbm    algorithmic bytes (n * W * 8 * G + 4 n + 32 per draw) against the measured copy peak (matters for 2..5 reads)

    tools/sweep_sizes.py homogeneous > gpurun_out/sweep_homogeneous.txt
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import ldpgta_b200 as L

# windows, draws per window: enough draws to fill the machine, short enough to finish in seconds
CFG = {2: (20000, 32), 3: (20000, 32), 4: (28738, 32), 5: (8000, 32), 6: (8000, 32), 7: (6000, 32), 8: (4000, 32), 9: (4000, 16),
       10: (4000, 8), 11: (2000, 8), 12: (2000, 4), 13: (1480, 2), 14: (1036, 2), 15: (592, 2), 16: (592, 2), 17: (296, 2), 18: (296, 1)}


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else 'homogeneous'
    sizes_n = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else list(range(2, 19))
    peaks = json.load(open(os.path.join(ROOT, 'profiles', 'INT_PEAKS.json')))
    hbm = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
    wordop, popc32 = peaks['wordop64_and_popc_add']['ops_per_s'], peaks['popc32']['ops_per_s']
    imad32, dfma = peaks['imad32']['ops_per_s'], peaks['dfma']['ops_per_s']
    dev = torch.device('cuda', 0)
    group_sizes, kind, prop = ([bench.N_HAP], L.HOMOGENEOUS, None) if model == 'homogeneous' else \
        ([2504, 2504], L.RECENT_ADMIXTURE, None) if model == 'recent' else ([2504, 2504], L.DISTANT_ADMIXTURE, [0.6, 0.4])
    words = sum((h + 63) // 64 for h in group_sizes)
    print('# %s, %s haplotypes, one B200; kernel-only medians' % (model, '+'.join(map(str, group_sizes))))
    for n in sizes_n:
        if model == 'distant' and n == 5:
            continue                                         # the reference has no likelihoods5 for distant admixture
        windows, subs = CFG[n]
        wins, reads = bench.workload_shape(0.2, 12345, min_reads=n + 2, max_windows=windows)
        eng = L.Engine(group_sizes)
        keep = []
        for g, nh in enumerate(group_sizes):
            m = bench.synth_masks(torch, dev, reads, nh, 12345 + 77 * g)
            eng.attach_read_masks_dev(g, m.data_ptr(), int(reads.sum()), keepalive=m)
        idx, _ = bench.plan_draws(reads, subs, n, 12346)
        d_idx = torch.from_numpy(idx).to(dev)
        d_out = torch.empty((len(idx), 4), dtype=torch.float64, device=dev)
        stream = torch.cuda.ExternalStream(eng.stream, device=dev)
        step = lambda: eng.likelihoods_uniform_dev(kind, n, d_idx.data_ptr(), len(idx), d_out.data_ptr(), proportions=prop)
        for _ in range(3):
            step()
        eng.synchronize()
        steps = 30 if n <= 8 else 7
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(stream)
        for k in range(steps):
            step()
            ev[k + 1].record(stream)
        eng.synchronize()
        ms = float(np.median([ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]))
        draws = len(idx)
        rate = draws / (ms * 1e-3)
        ops = ((1 << n) - 1) * words
        t_count = ops * (4.0 / 3.0) / popc32                 # carry-save: three words, two POPC.64 = four POPC.32
        if n >= 5:
            terms = (3 ** (n - 5) - 1) / 2 + 1
            run = 8 if n >= 15 else 4 if n >= 12 else 1
            t_poly = terms * (81 + 32 / run if model == 'homogeneous' else 4 * 81 + 4 * 32 / run if model == 'recent' else 4 * 81 + 6 * 32 / run) / imad32
        else:
            t_poly = 0.0
        two = (t_count + t_poly) * rate                      # fraction of the two-term ceiling
        bytes_per_draw = n * words * 8 + 4 * n + 32
        print('%s n=%2d  %-18s draws/step %7d  ms %9.4f  draws/s %.3e  popc %.3f  two-term %.3f (poly share of ceiling %.2f)  hbm %.3f'
              % (model, n, bench.kernel_name(n, model), draws, ms, rate, ops * rate / wordop, two, t_poly / (t_count + t_poly), bytes_per_draw * rate / 1e9 / hbm))
        sys.stdout.flush()
        eng.close()
        del keep, d_idx, d_out
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
