#!/usr/bin/env python
""" Where a resident step of configs[1] goes, measured in the pipeline (warm L2, back-to-back launches) rather than under
ncu: median step with no inner events, with events around the draw kernel, with events around the reductions.
usage: python tools/step_parts.py [--subsamples S] [--windows W] [--list-fraction F] """
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--subsamples', type=int, default=32)
    ap.add_argument('--windows', type=int, default=0)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--list-fraction', type=float, default=None, help='windows as read lists keeping this fraction of their rows (bench.Leg)')
    a = ap.parse_args()
    import torch
    import ldpgta_b200 as L
    device = torch.device('cuda:0')
    wins, reads = B.workload_shape(0.1, 12345, min_reads=6, max_windows=a.windows)
    leg = B.Leg(L, torch, device, 0, [B.N_HAP], L.HOMOGENEOUS, reads, 4, a.subsamples, 12345, None, segments=B.chromosome_segments(wins, len(reads)),
                list_fraction=a.list_fraction)
    eng = leg.eng

    def loop(mode):
        for _ in range(20):
            leg.step()
        eng.synchronize()
        eng.kernel_timing(mode)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        ev[0].record(leg.stream)
        for k in range(a.steps):
            leg.step()
            ev[k + 1].record(leg.stream)
        eng.synchronize()
        inner = eng.kernel_times() if mode else np.zeros(1)
        eng.kernel_timing(False)
        return float(np.median([ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps)])), float(np.median(inner)), ev[0].elapsed_time(ev[-1]) / a.steps
    out = {}
    for name, mode in (('no_inner_events', 0), ('events_around_draw_kernel', 1), ('events_around_reductions', 2), ('events_around_window_statistics', 3), ('events_around_sampler', 4), ('no_inner_events_again', 0)):
        step, inner, mean = loop(mode)
        out[name] = {'median_step_ms': step, 'mean_step_ms': mean, 'median_bracket_ms': inner}
    out['draws_per_step'] = leg.n_draws
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
