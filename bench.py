#!/usr/bin/env python3
"""
bench.py -- bootstrapped window-likelihoods per second (BASELINE.json metric).

Workload (BASELINE.json configs[1]): one synthetic embryo, 22 autosomes at 0.1x, 100 kbp windows (about 28.7k),
~150 informative reads per window, 5008-haplotype LD-structured panel, 32 bootstrap draws of 4 reads per window (the
CLI defaults -w 100000 -s 32 -m 6 -M 4), all four hypotheses (monosomy, disomy, SPH, BPH) per draw, LLR statistics per
window and per chromosome.  A "step" is one pass of the hot path over the embryo: pick_reads for every (window, draw),
the likelihoods of every draw, LLRs, window mean / variance and the 22 chromosome summaries
(ANEUPLOIDY_TEST.py:226-325).  With N GPUs every rank owns one such embryo (weak scaling, no data-path collective:
(embryo, chromosome, window, draw) units are independent).

  value      draws/s of the device-resident pipeline (ldpgta_bootstrap_stats_dev: sampler + draw kernel + reductions;
             read masks, windows and every intermediate in HBM), CUDA events on the engine's stream, max over ranks;
             `sustained` repeats it for >= 0.3 s and reports the median step
  e2e        the same step through the public host API with HOST buffers, copies inside the timed region:
               stats   Engine.bootstrap_stats -> ldpgta_bootstrap_stats: the seed goes up, the statistics come down
                       (80 B per window + the chromosome summaries) -- this is e2e.value
               full    the same plus the per-draw likelihoods (32 B per draw) streamed down under the kernels
               replay  Engine.replay_stats -> ldpgta_replay_stats: draws chosen on the host (the reference's
                       random.sample stream) go up (16 B per draw), statistics come down
  roofline   dominant kernel k_small<4,8,10>: algorithmic bytes / launch time (CUDA events around the kernel inside the
             timed steps) vs the measured HBM copy peak; int_roofline: algorithmic AND+POPC word-ops vs the POPC peak
  secondary  short legs for the other BASELINE configs: configs[2] (16 reads per draw), configs[3] (recent admixture,
             two populations) and 1000 subsamples per window, each with ms per step, draws/s and roofline fraction
  strong / batch96 (N > 1 and, batch96, N = 1 too): ONE embryo's windows dealt to the ranks with the chromosome
             partial sums all-reduced inside the timed step; configs[4]: 96 embryos x 1000 subsamples dealt to the ranks
  cpu_baseline  the reference's own model class (oracle/_ref, unmodified) on one host core on a bounded sample of the
             same windows and draws -- also the in-run parity check; the oracle port's rate beside it

--impl reference times the reference's CPU loop (ANEUPLOIDY_TEST.py:277-286 on oracle/_ref's classes; the oracle port
when oracle/_ref is absent) on all host cores (one process per core) on bounded samples of the same workload.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'oracle')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# hg38 autosome lengths (the table MIX_HAPLOIDS.chr_length uses, MIX_HAPLOIDS.py:31-39)
AUTOSOME_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                    133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                    58617616, 64444167, 46709983, 50818468]
N_HAP = 5008
WINDOW = 100000
FOUNDERS = 64


def workload_shape(depth, seed, min_reads=6, max_windows=0):
    """ windows per chromosome and informative reads per window (~1500 * depth). """
    rng = np.random.default_rng(seed)
    wins = [L // WINDOW for L in AUTOSOME_LENGTHS]
    n_windows = int(sum(wins))
    reads = np.maximum(rng.poisson(1500.0 * depth, size=n_windows), min_reads).astype(np.int64)
    if max_windows:
        reads = reads[:max_windows]
    return wins, reads


def synth_masks(torch, device, reads_per_window, n_hap, seed, window_slice=None):
    """ Read masks [n_reads, padded_words] (int64 bit patterns) for an LD-structured panel:
    inside a window every haplotype copies one of 64 founders (the assignment drifts from
    window to window); a read matches a founder with a per-allele frequency drawn from a skewed
    spectrum and carries 1-3 alleles, so its mask over the panel is a 64-bit founder word
    expanded through the window's haplotype->founder map.  Deep intersections stay non-empty. """
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    words = (n_hap + 63) // 64
    wp = (words + 1) & ~1
    sl = window_slice or slice(0, len(reads_per_window))
    rpw = np.asarray(reads_per_window[sl])
    starts = np.concatenate([[0], np.cumsum(rpw)])
    total = int(starts[-1])
    masks = torch.zeros((total, wp), dtype=torch.int64, device=device)
    assign = torch.randint(0, FOUNDERS, (n_hap,), generator=g, device=device)
    shifts = torch.arange(64, device=device, dtype=torch.int64)
    chunk = max(1, min(32, int(20000 // max(1, rpw.max()))))
    for w0 in range(0, len(rpw), chunk):
        w1 = min(w0 + chunk, len(rpw))
        nr = int(starts[w1] - starts[w0])
        amap = torch.empty((w1 - w0, n_hap), dtype=torch.int64, device=device)
        for k in range(w1 - w0):                        # founder map drifts: 5% of haplotypes switch per window
            sw = torch.rand((n_hap,), generator=g, device=device) < 0.05
            assign = torch.where(sw, torch.randint(0, FOUNDERS, (n_hap,), generator=g, device=device), assign)
            amap[k] = assign
        win_of_read = torch.repeat_interleave(torch.arange(w1 - w0, device=device),
                                              torch.as_tensor(rpw[w0:w1], device=device))
        # founder word of each read: AND over its 1..3 alleles of Bernoulli(u) founder bits, u skewed
        k_alleles = 1 + (torch.rand((nr,), generator=g, device=device) > 0.68).long() + (torch.rand((nr,), generator=g, device=device) > 0.93).long()
        fbits = torch.ones((nr, FOUNDERS), dtype=torch.bool, device=device)
        for a in range(3):
            u = torch.rand((nr, 1), generator=g, device=device) ** 0.6
            draw = torch.rand((nr, FOUNDERS), generator=g, device=device) < u
            fbits &= draw | (k_alleles[:, None] <= a)
        fword = (fbits.long() << shifts).sum(dim=1)                                  # wraps into the sign bit as intended
        bits = (fword[:, None] >> amap[win_of_read]) & 1                             # [nr, n_hap]
        padded = torch.zeros((nr, wp * 64), dtype=torch.int64, device=device)
        padded[:, :n_hap] = bits
        masks[int(starts[w0]):int(starts[w1])] = (padded.view(nr, wp, 64) << shifts).sum(dim=2)
    return masks


def plan_draws(reads_per_window, subsamples, max_reads, seed):
    """ Host-side draws (the replay leg and the reference arm): for every window effN = min(C(R, max_reads), subsamples)
    draws (ANEUPLOIDY_TEST.py:235-248) of n = min(R-1, max_reads) distinct reads (:231), as global read indices. """
    rng = np.random.default_rng(seed)
    R = np.asarray(reads_per_window)
    assert R.min() > max_reads, 'uniform draw size needs R > max_reads in every window'
    n = max_reads
    starts = np.concatenate([[0], np.cumsum(R)])[:-1]
    win = np.repeat(np.arange(len(R)), subsamples)
    Rd = R[win]
    # Floyd's algorithm, vectorised over draws: distinct samples without replacement
    chosen = np.empty((len(win), n), dtype=np.int64)
    for j in range(n):
        upper = Rd - n + j                                   # candidate in [0, upper]
        t = (rng.random(len(win)) * (upper + 1)).astype(np.int64)
        dup = (chosen[:, :j] == t[:, None]).any(axis=1) if j else np.zeros(len(win), dtype=bool)
        chosen[:, j] = np.where(dup, upper, t)
    perm = np.argsort(rng.random((len(win), n)), axis=1)     # draws are unordered samples
    chosen = np.take_along_axis(chosen, perm, axis=1)
    idx = (chosen + starts[win][:, None]).astype(np.int32)
    return idx, win


class ClockSampler:
    """ nvidia-smi clocks / throttle reasons during the timed region. """
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ---- the CPU arm: the reference's own classes (oracle/_ref) or the oracle port --------------------------------------
def cpu_model_class(prefer_reference=True):
    """ (class with the signature of HOMOGENOUES_MODELS.homogeneous, kind, popcount description) """
    if prefer_reference:
        try:
            import make_ref
            if make_ref.available():
                # loaded by file path under a private name: the bare name HOMOGENOUES_MODELS may already belong to the
                # GPU drop-in (ldpgta_b200 registers it), and the CPU arm must never time that
                import importlib.util
                name = '_ldpgta_reference_HOMOGENOUES_MODELS'
                ref = sys.modules.get(name)
                if ref is None:
                    spec = importlib.util.spec_from_file_location(name, os.path.join(make_ref.DEST, 'HOMOGENOUES_MODELS.py'))
                    ref = importlib.util.module_from_spec(spec)
                    sys.modules[name] = ref
                    with contextlib.redirect_stdout(io.StringIO()):
                        spec.loader.exec_module(ref)
                assert ref.__file__.startswith(make_ref.DEST) and not hasattr(ref.homogeneous, '_engine')
                pc = 'gmpy2.popcount' if type(ref.popcount).__name__ != 'popcount_lk' else 'popcount_lk (byte LUT; gmpy2 is not installed)'
                return ref.homogeneous, 'reference', pc
        except Exception as e:                                  # noqa: the port below is the documented fallback
            sys.stderr.write('bench: oracle/_ref unusable (%s), timing the oracle port\n' % e)
    import ldpgta_oracle as O
    return O.Homogeneous, 'port', 'int.bit_count (reference: gmpy2.popcount or byte LUT)'


def cpu_window_model(cls, masks_rows, first_read, n_hap):
    """ The model object for one window (built once, outside the timed loop, like homogeneous.__init__ in the
    reference's bootstrap()).  Reads are single-allele haplotypes (read_index, 'A'). """
    words = (n_hap + 63) // 64
    ints = [int.from_bytes(r[:words].tobytes(), 'little') for r in masks_rows]
    leg = [('chrS', first_read + i, 'C', 'A') for i in range(len(ints))]
    obs = [(first_read + i, 'read%d' % (first_read + i), 'A') for i in range(len(ints))]
    with contextlib.redirect_stdout(io.StringIO()):
        model = cls(obs, leg, {'ALL': ints}, {'ALL': n_hap}, {})
    reads = {first_read + i: [(first_read + i, 'A')] for i in range(len(ints))}
    return model, reads


def cpu_loop(model, reads, draws):
    """ The timed unit: the bootstrap inner loop (ANEUPLOIDY_TEST.py:282-286 minus the progress bar). """
    t0 = time.perf_counter()
    out = [tuple(model.get_likelihoods(tuple(reads[int(r)] for r in d))) for d in draws]
    return time.perf_counter() - t0, out


_W = {}


def _worker_init(masks, starts, idx, win, n_hap):
    cls, _, _ = cpu_model_class()
    _W.update(masks=masks, starts=starts, idx=idx, win=win, n_hap=n_hap, models={}, cls=cls)


def _worker_step(windows):
    n = 0
    for w in windows:
        if w not in _W['models']:
            r0, r1 = int(_W['starts'][w]), int(_W['starts'][w + 1])
            _W['models'][w] = cpu_window_model(_W['cls'], _W['masks'][r0:r1], r0, _W['n_hap']) + (_W['idx'][_W['win'] == w],)
        model, reads, draws = _W['models'][w]
        cpu_loop(model, reads, draws)
        n += len(draws)
    return n


def run_reference(args):
    """ --impl reference: the reference's CPU loop on all host cores (one process per core, windows dealt
    round-robin; model objects are built once per window outside the timed steps). """
    import multiprocessing as mp
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    _, kind, pc = cpu_model_class()
    wins, reads = workload_shape(args.depth, args.seed, min_reads=args.max_reads + 2, max_windows=args.windows)
    per_step_windows = min(max(cores * 32, 256), len(reads))  # bounded sample, the same windows every step
    need = per_step_windows
    masks = synth_masks(torch, 'cpu', reads, N_HAP, args.seed, window_slice=slice(0, need)).numpy().view(np.uint64)
    idx, win = plan_draws(reads[:need], args.subsamples, args.max_reads, args.seed + 1)
    starts = np.concatenate([[0], np.cumsum(reads[:need])])
    pool = mp.get_context('fork').Pool(cores, initializer=_worker_init, initargs=(masks, starts, idx, win, N_HAP))
    chunks = [list(range(i, per_step_windows, cores)) for i in range(cores)]
    chunks = [c for c in chunks if c]

    def step():
        t0 = time.perf_counter()
        n = sum(pool.map(_worker_step, chunks, chunksize=1))
        return n, time.perf_counter() - t0
    step()                                                     # builds the per-window models in the workers
    for _ in range(args.warmup):
        step()
    n_total, t_total = 0, 0.0
    for _ in range(args.steps):
        n, t = step()
        n_total += n; t_total += t
    pool.close()
    value = n_total / t_total
    sample = '%d windows x %d draws per step (%d steps) of the %d-window embryo' % (per_step_windows, args.subsamples, args.steps, len(reads))
    line = {'impl': 'reference', 'metric': 'bootstrapped window-likelihoods/sec', 'value': value, 'unit': 'window-likelihoods/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64 popcount + f64', 'data': 'synthetic',
            'config': workload_config(args, len(reads), int(reads.sum())),
            'cpu_baseline': {'value': value, 'unit': 'window-likelihoods/s', 'cores': cores, 'kind': kind, 'sample': sample,
                             'python': sys.version.split()[0], 'popcount': pc},
            'e2e': {'value': value, 'unit': 'window-likelihoods/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)
    return 0


def kernel_name(n, model, subsamples=32):
    """ the kernel ldpgta_api.cu::run_uniform / bootstrap_api.cu dispatch draws of n reads to """
    if n == 4 and model == 'homogeneous' and subsamples >= 256:
        return 'k_small_win (window-resident masks)'
    if n <= 4:
        return 'k_small<%d,8,10>' % n
    if n == 5 and model == 'homogeneous':
        return 'k_mid<5,5>'
    if n <= 12 and not (model == 'distant' and n == 5):        # (distant admixture of three or more groups: up to 10)
        return 'k_lanes<%d,%d>' % (n, {5: 3, 6: 3, 7: 3, 8: 4, 9: 4}.get(n, 5))
    return 'k_general (n=%d)' % n


def workload_config(args, n_windows, n_reads):
    tag = 'configs[1]' if (args.max_reads == 4 and args.depth == 0.1 and not args.windows and args.model == 'homogeneous') else 'exploration run'
    return {'workload': '%s: 22-autosome synthetic embryo, depth %.2fx, homogeneous model, monosomy/disomy/SPH/BPH' % (tag, args.depth),
            'haplotypes': N_HAP, 'windows_per_gpu': n_windows, 'reads_per_gpu': n_reads, 'window_size': WINDOW,
            'subsamples': args.subsamples, 'max_reads': args.max_reads, 'min_reads': 6,
            'panel': 'LD-structured (64-founder mosaic), synthetic', 'l2': 'inputs larger than L2 (read masks %.2f GB per GPU)' % (n_reads * 640 / 1e9),
            'parallelism': 'one embryo per GPU, windows independent, no data-path collective'}


_JSON_OUT = None


def claim_stdout():
    """ The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to
    stdout), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor. """
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')


def emit(line):
    _JSON_OUT.write(json.dumps(line) + '\n')
    _JSON_OUT.flush()


def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    int_peak, int_src = 2.3e12, 'assumed 148 SM x 16 POPC.32/clk x 1.965 GHz / 2'
    try:
        ip = json.load(open(os.path.join(ROOT, 'profiles', 'INT_PEAKS.json')))
        int_peak = float(ip['wordop64_and_popc_add']['ops_per_s'])
        int_src = 'measured (profiles/INT_PEAKS.json, tools/int_peaks.cu: AND+POPC+ADD on 64-bit words)'
    except Exception:
        pass
    return hbm_peak, peak_src, int_peak, int_src


def chromosome_segments(wins, n_windows):
    """ window offsets of the 22 chromosomes, clipped to the first n_windows windows """
    seg = np.minimum(np.concatenate([[0], np.cumsum(wins)]), n_windows).astype(np.int64)
    return seg[:int(np.searchsorted(seg, n_windows, side='left')) + 1]


class Leg:
    """ One resident workload: masks + plan on an Engine, timed through ldpgta_bootstrap_stats_dev. """

    def __init__(self, L, torch, device, local, group_sizes, kind, reads, n, subsamples, seed, proportions=None, segments=None,
                 seg_cols=None, seg_first=None, list_fraction=None):
        self.L, self.torch, self.kind, self.n, self.proportions = L, torch, kind, n, proportions
        self.reads, self.n_windows, self.n_reads = reads, len(reads), int(reads.sum())
        self.eng = L.Engine(group_sizes, device=local)
        self.masks = []
        for g, nh in enumerate(group_sizes):
            m = synth_masks(torch, device, reads, nh, seed + 77 * g)
            assert m.shape[1] == self.eng.padded_words[g]
            self.eng.attach_read_masks_dev(g, m.data_ptr(), self.n_reads, keepalive=m)
            self.masks.append(m)
        self.woff = np.concatenate([[0], np.cumsum(reads)]).astype(np.int64)
        self.doff = np.arange(0, (self.n_windows + 1) * subsamples, subsamples, dtype=np.int64)
        self.n_draws = int(self.doff[-1])
        wreads = None
        if list_fraction is not None:
            # windows as read lists, the way packed.py hands them over: a window keeps list_fraction of its run of rows (reads
            # below min_score leave gaps), at least n + 2 of them
            rng = np.random.default_rng(seed + 4242)
            keep = rng.random(self.n_reads) < list_fraction
            first = np.arange(self.n_reads) - np.repeat(self.woff[:-1], reads)
            keep |= first < n + 2
            wreads = np.flatnonzero(keep).astype(np.int32)
            self.woff = np.concatenate([[0], np.cumsum(np.add.reduceat(keep, self.woff[:-1]))]).astype(np.int64)
        self.info = self.eng.set_windows(self.woff, wreads, self.doff, n, segment_offsets=segments, segment_cols=seg_cols, segment_first=seg_first)
        self.stream = torch.cuda.ExternalStream(self.eng.stream, device=device)
        self.words = sum((nh + 63) // 64 for nh in group_sizes)
        self.seed = 0

    def step(self, partials_ptr=None):
        self.seed += 1
        self.eng.bootstrap_stats_dev(self.kind, self.seed, proportions=self.proportions, partials_dev_ptr=partials_ptr)

    def timed(self, steps, warmup, min_seconds=0.0):
        """ per-step milliseconds (CUDA events on the engine's stream) and per-step draw-kernel milliseconds """
        torch = self.torch
        for _ in range(warmup):
            self.step()
        self.eng.synchronize()
        if min_seconds:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(self.stream); self.step(); ev1.record(self.stream)
            self.eng.synchronize()
            steps = max(steps, min(20000, int(min_seconds * 1e3 / max(ev0.elapsed_time(ev1), 1e-3)) + 1))
        self.eng.kernel_timing(True, stride=4)                    # the draw kernel of every fourth step is bracketed (a pair of events costs ~5 us)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(self.stream)
        for k in range(steps):
            self.step()
            ev[k + 1].record(self.stream)
        self.eng.synchronize()
        kms = self.eng.kernel_times()
        self.eng.kernel_timing(False)
        return np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]), kms, ev[0].elapsed_time(ev[-1])

    def close(self):
        self.eng.close()
        self.masks = []


def secondary_legs(L, torch, device, local, args, hbm_peak, int_peak):
    """ short legs for the other BASELINE configs on rank 0 (resident pipeline, device draws, statistics included) """
    out = []

    def run(name, group_sizes, kind, depth, n, subsamples, windows, steps, proportions=None, model='homogeneous', list_fraction=None):
        wins, reads = workload_shape(depth, args.seed + 5, min_reads=n + 2, max_windows=windows)
        leg = Leg(L, torch, device, local, group_sizes, kind, reads, n, subsamples, args.seed + 9, proportions,
                  segments=chromosome_segments(wins, len(reads)), list_fraction=list_fraction)
        ms, kms, _ = leg.timed(steps, 3, min_seconds=0.2)
        med, kmed = float(np.median(ms)), float(np.median(kms))
        wordops = ((1 << n) - 1) * leg.words
        bytes_per_draw = n * leg.words * 8 + n * 4 + 32
        out.append({'workload': name, 'kernel': kernel_name(n, model, subsamples), 'windows': leg.n_windows, 'reads': leg.n_reads, 'draws_per_step': leg.n_draws,
                    'reads_per_draw': n, 'subsamples': subsamples, 'steps': len(ms), 'ms_per_step': med, 'draw_kernel_ms': kmed,
                    'value': leg.n_draws / (med * 1e-3), 'unit': 'window-likelihoods/s',
                    'hbm_frac': bytes_per_draw * leg.n_draws / (kmed * 1e-3) / 1e9 / hbm_peak,
                    'int_frac': wordops * leg.n_draws / (kmed * 1e-3) / int_peak})
        leg.close()
        torch.cuda.empty_cache()

    run('configs[2]: 1x coverage, 16 reads per draw (2^16 subsets), homogeneous, 1024 windows x 8 draws', [N_HAP], L.HOMOGENEOUS, 1.0, 16, 8, 1024, 5)
    run('configs[3]: recent admixture, two populations of 2504 haplotypes, 22 autosomes at 0.1x, 4 reads per draw, 32 draws per window',
        [N_HAP // 2, N_HAP // 2], L.RECENT_ADMIXTURE, 0.1, 4, 32, 0, 20, model='recent')
    run('configs[4] shape: 1000 subsamples per window, 4 reads per draw, homogeneous, 2048 windows at 0.1x', [N_HAP], L.HOMOGENEOUS, 0.1, 4, 1000, 2048, 5)
    run('configs[4] shape, windows as read lists (85 % of each window\'s rows, as packed.py hands them over): 1000 subsamples per window, '
        '4 reads per draw, homogeneous, 2048 windows at 0.1x', [N_HAP], L.HOMOGENEOUS, 0.1, 4, 1000, 2048, 5, list_fraction=0.85)
    return out


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--depth', type=float, default=0.1)
    ap.add_argument('--subsamples', type=int, default=32)
    ap.add_argument('--max-reads', type=int, default=4)
    ap.add_argument('--seed', type=int, default=12345)
    ap.add_argument('--model', default='homogeneous', choices=['homogeneous', 'recent', 'distant'], help='recent / distant = two-population exploration runs')
    ap.add_argument('--windows', type=int, default=0, help='limit the number of windows (exploration runs; 0 = whole genome)')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='budget of the single-core CPU baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the secondary legs (configs[2], configs[3], 1000 subsamples)')
    ap.add_argument('--no-batch96', action='store_true', help='skip the configs[4] leg')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import ldpgta_b200 as L

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the likelihood engine has no CPU path')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_peak, peak_src, int_peak, int_src = load_peaks()

    # ---- workload: one embryo per rank ------------------------------------------------------
    wins, reads = workload_shape(args.depth, args.seed + 1000 * rank, min_reads=args.max_reads + 2, max_windows=args.windows)
    n = args.max_reads
    proportions = None
    if args.model == 'recent':                       # exploration: configs[3], two populations of 2504 haplotypes
        group_sizes, kind = [N_HAP // 2, N_HAP // 2], L.RECENT_ADMIXTURE
    elif args.model == 'distant':                    # exploration: the same panel with ancestral proportions 0.6 / 0.4
        group_sizes, kind, proportions = [N_HAP // 2, N_HAP // 2], L.DISTANT_ADMIXTURE, [0.6, 0.4]
    else:
        group_sizes, kind = [N_HAP], L.HOMOGENEOUS
    segments = chromosome_segments(wins, len(reads))
    leg = Leg(L, torch, device, local, group_sizes, kind, reads, n, args.subsamples, args.seed + 1000 * rank, proportions, segments=segments)
    eng, n_windows, n_reads, n_draws, n_seg = leg.eng, leg.n_windows, leg.n_reads, leg.n_draws, len(segments) - 1

    # ---- value: the resident pipeline, K steps ------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.5:               # untimed: let the SM clock leave idle before warm-up
        leg.step()
        eng.synchronize()
    for _ in range(args.warmup):
        leg.step()
    eng.synchronize()
    barrier()
    launches0 = eng.launch_count
    step_ms, kernel_ms, total_ms = leg.timed(args.steps, 0)
    barrier()
    launches = eng.launch_count - launches0
    total_ms = max_over_ranks(total_ms)
    value = world * n_draws * args.steps / (total_ms * 1e-3)
    # the same loop for at least 0.3 s: median step and median draw-kernel time (8 ms regions wobble by +-5 %)
    sus_ms, sus_kms, _ = leg.timed(args.steps, 0, min_seconds=0.3)
    sustained = {'steps': int(len(sus_ms)), 'median_ms_per_step': float(np.median(sus_ms)), 'mean_ms_per_step': float(np.mean(sus_ms)),
                 'p10_ms': float(np.percentile(sus_ms, 10)), 'p90_ms': float(np.percentile(sus_ms, 90)),
                 'median_draw_kernel_ms': float(np.median(sus_kms)),
                 'value_at_median': world * n_draws / (max_over_ranks(float(np.median(sus_ms))) * 1e-3)}

    # ---- e2e: host buffers through the public API ----------------------------------------------
    ncp = leg.info['ncp']
    pin = {'mean_var': L.PinnedArray((5, n_windows, 2), np.float64), 'stats': L.PinnedArray((n_seg, 5, 3), np.float64),
           'lik': L.PinnedArray((n_draws, 4), np.float64)}
    out_stats = {'mean_var': pin['mean_var'].array, 'stats': pin['stats'].array}
    out_full = dict(out_stats, lik=pin['lik'].array)
    idx_host, win_of_draw = plan_draws(reads, args.subsamples, n, args.seed + 1 + 1000 * rank)
    pin_idx = L.PinnedArray((n_draws, n), np.int32)
    pin_idx.array[:] = idx_host
    seeds = iter(range(10 ** 6, 10 ** 9))

    def e2e_stats():
        eng.bootstrap_stats(kind, next(seeds), proportions=proportions, out=out_stats)

    def e2e_full():
        eng.bootstrap_stats(kind, next(seeds), proportions=proportions, out=out_full)

    def e2e_replay():
        eng.replay_stats(kind, pin_idx.array, proportions=proportions, out=out_stats)

    def time_host(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        barrier()
        return max_over_ranks(time.perf_counter() - t0)

    e2e_s = {name: time_host(fn) for name, fn in (('stats', e2e_stats), ('full', e2e_full), ('replay', e2e_replay))}
    clocks = sampler.stop()
    stats_b = int(out_stats['mean_var'].nbytes + out_stats['stats'].nbytes)
    e2e_modes = {
        'stats': {'value': world * n_draws * args.steps / e2e_s['stats'], 'ms_per_step': 1e3 * e2e_s['stats'] / args.steps,
                  'h2d_bytes_per_step': 8, 'd2h_bytes_per_step': stats_b,
                  'api': 'Engine.bootstrap_stats -> ldpgta_bootstrap_stats: draws on the device; window mean/variance + chromosome summaries come down'},
        'full': {'value': world * n_draws * args.steps / e2e_s['full'], 'ms_per_step': 1e3 * e2e_s['full'] / args.steps,
                 'h2d_bytes_per_step': 8, 'd2h_bytes_per_step': stats_b + int(pin['lik'].array.nbytes),
                 'api': 'the same call, also returning the per-draw likelihoods (32 B per draw)'},
        'replay': {'value': world * n_draws * args.steps / e2e_s['replay'], 'ms_per_step': 1e3 * e2e_s['replay'] / args.steps,
                   'h2d_bytes_per_step': int(pin_idx.array.nbytes), 'd2h_bytes_per_step': stats_b,
                   'api': 'Engine.replay_stats -> ldpgta_replay_stats: host-chosen draws go up (16 B per draw), statistics come down'}}

    # ---- consistency: replayed device draws give the same likelihoods and statistics through every route ----
    chk = eng.bootstrap_stats(kind, 4242, proportions=proportions, want=('lik', 'draws', 'mean_var', 'stats'))
    rep = eng.replay_stats(kind, chk['draws'], proportions=proportions, want=('lik', 'mean_var', 'stats'))
    assert np.array_equal(chk['lik'], rep['lik']) and np.array_equal(chk['stats'], rep['stats']), 'device-drawn and replayed results differ'
    d_chk = torch.from_numpy(chk['draws'].reshape(n_draws, n)).to(device)
    d_out = torch.empty((n_draws, 4), dtype=torch.float64, device=device)
    eng.likelihoods_uniform_dev(kind, n, d_chk.data_ptr(), n_draws, d_out.data_ptr(), proportions=proportions)
    eng.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), chk['lik']), 'pipeline and plain batch call differ'

    # ---- multi-GPU legs --------------------------------------------------------------------------
    strong = None
    if dist is not None:
        strong = strong_scaling_leg(L, torch, dist, device, local, rank, world, args, kind, group_sizes, proportions, barrier, max_over_ranks)
    batch96 = None
    if not args.no_batch96 and args.model == 'homogeneous' and not args.windows:
        batch96 = batch96_leg(L, torch, dist, device, rank, world, leg, barrier, max_over_ranks, segments)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- rooflines of the dominant kernel ------------------------------------------------------
    words = leg.words
    bytes_per_draw = n * words * 8 + n * 4 + 32               # n*W*8 mask bytes + indices in, 32 B out
    wordops_per_draw = ((1 << n) - 1) * words
    k_ms = float(np.mean(kernel_ms))
    achieved_gbs = bytes_per_draw * n_draws / (k_ms * 1e-3) / 1e9
    achieved_wordops = wordops_per_draw * n_draws / (k_ms * 1e-3)
    traffic = None                                            # DRAM bytes per launch from the committed ncu capture of this kernel
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'))).get('k_small<%d,8,10>' % n)
        if t and n <= 4:
            traffic = t['dram_bytes_per_launch'] * n_draws / t['draws_per_launch']
    except Exception:
        pass

    # ---- CPU baseline: the reference's class on a bounded sample, one core; doubles as the parity check -------
    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline and world == 1 and args.model == 'homogeneous':     # rank 0 at N=1 only
        masks = leg.masks[0]
        draws = chk['draws'].reshape(n_draws, n)
        starts = leg.woff

        def sample(cls, budget, wall):
            spent, n_eval, w, t_wall = 0.0, 0, 0, time.perf_counter()
            ref_rows = []
            while spent < budget and w < n_windows and time.perf_counter() - t_wall < wall:
                r0, r1 = int(starts[w]), int(starts[w + 1])
                model, rd = cpu_window_model(cls, masks[r0:r1].cpu().numpy().view(np.uint64), r0, N_HAP)
                t, vals = cpu_loop(model, rd, draws[w * args.subsamples:(w + 1) * args.subsamples])
                ref_rows.append(np.array(vals, dtype=np.float64).reshape(-1, 4))
                spent += t; n_eval += args.subsamples; w += 1
            return spent, n_eval, w, np.concatenate(ref_rows)

        cls, kind_cpu, pc = cpu_model_class()
        spent, n_eval, w, ref_v = sample(cls, args.cpu_seconds, 4 * args.cpu_seconds + 20)
        got_v = chk['lik'][:n_eval]
        parity = float(np.max(np.abs(ref_v - got_v) / np.maximum(np.abs(ref_v), 1e-300)))
        cpu_baseline = {'value': n_eval / spent, 'unit': 'window-likelihoods/s', 'cores': 1, 'kind': kind_cpu,
                        'sample': 'first %d windows x %d draws (%d draws, %.1f s) of the same embryo, the draws the device made' % (w, args.subsamples, n_eval, spent),
                        'python': sys.version.split()[0], 'popcount': pc, 'host_cores_available': os.cpu_count()}
        if kind_cpu == 'reference':                            # the oracle port's rate beside it (int.bit_count instead of the byte LUT)
            pcls, _, ppc = cpu_model_class(prefer_reference=False)
            ps, pn, pw, pref = sample(pcls, min(3.0, args.cpu_seconds), 20)
            m = min(len(pref), len(ref_v))
            cpu_baseline['port'] = {'value': pn / ps, 'popcount': ppc, 'sample': 'first %d windows (%.1f s)' % (pw, ps),
                                    'max_rel_diff_vs_reference': float(np.max(np.abs(pref[:m] - ref_v[:m]) / np.maximum(np.abs(ref_v[:m]), 1e-300)))}

    secondary = None
    if not args.no_secondary and args.model == 'homogeneous' and not args.windows:
        leg.close()
        del leg, eng
        torch.cuda.empty_cache()
        try:
            secondary = secondary_legs(L, torch, device, local, args, hbm_peak, int_peak)
        except Exception as e:                                  # the headline line must survive a failing side leg
            secondary = [{'error': '%s: %s' % (type(e).__name__, e)}]

    line = {
        'metric': 'bootstrapped window-likelihoods/sec', 'value': value, 'unit': 'window-likelihoods/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u64 and/popcount + f64 polynomials', 'data': 'synthetic',
        'config': workload_config(args, n_windows, n_reads),
        'draws_per_step_per_gpu': int(n_draws),
        'step': 'device-side draws (Philox + Floyd) -> likelihoods -> LLRs -> window mean/variance -> %d chromosome summaries' % n_seg,
        'sustained': sustained,
        'e2e': dict(e2e_modes['stats'], unit='window-likelihoods/s', mode='stats', **{k: v for k, v in e2e_modes.items()}),
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': kernel_name(n, args.model), 'achieved': achieved_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': achieved_gbs / hbm_peak, 'traffic': traffic, 'peak_source': peak_src,
                     'algorithmic_bytes_per_draw': bytes_per_draw, 'kernel_ms': k_ms,
                     'kernel_share_of_step': k_ms / (total_ms / args.steps)},
        'int_roofline': {'bound': 'popc', 'achieved': achieved_wordops, 'peak': int_peak, 'unit': 'word AND+POPC/s',
                         'frac': achieved_wordops / int_peak, 'peak_source': int_src, 'wordops_per_draw': wordops_per_draw},
        'cpu_baseline': cpu_baseline,
        'parity_max_rel_err_vs_oracle': parity,
        'clocks': clocks,
    }
    if secondary is not None:
        line['secondary'] = secondary
    if strong is not None:
        line['strong'] = strong
    if batch96 is not None:
        line['batch96'] = batch96
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def strong_scaling_leg(L, torch, dist, device, local, rank, world, args, kind, group_sizes, proportions, barrier, max_over_ranks):
    """ ONE configs[1] embryo: its windows dealt to the ranks as contiguous ranges of equal cost
    (sharding.partition_windows), every rank computing the chromosome partial sums of its windows; ONE all_reduce of
    the [22][5][35] partials and the finish kernel inside the timed step, everything device-resident. """
    from ldpgta_b200 import sharding
    wins, reads = workload_shape(args.depth, args.seed, min_reads=args.max_reads + 2)
    n, subs = args.max_reads, args.subsamples
    seg = chromosome_segments(wins, len(reads))
    words = sum((nh + 63) // 64 for nh in group_sizes)
    costs = [sharding.draw_cost(int(R), 6, n, subs, words, len(group_sizes)) for R in reads]
    lo, hi = sharding.partition_windows(costs, world)[rank]
    my_seg = np.clip(seg, lo, hi) - lo
    n_seg = len(seg) - 1
    cols = np.full(n_seg, subs, dtype=np.int32)                  # every window of the embryo has `subs` draws
    leg = Leg(L, torch, device, local, group_sizes, kind, reads[lo:hi], n, subs, args.seed + 31 * rank, proportions, segments=my_seg,
              seg_cols=cols, seg_first=cols)
    ncp = leg.info['ncp']
    partials = torch.zeros((n_seg, 5, ncp + 3), dtype=torch.float64, device=device)
    stats = torch.zeros((n_seg, 5, 3), dtype=torch.float64, device=device)
    steps, warm = args.steps, args.warmup
    peer_ok = sharding.connect_peers(leg.eng, partials.numel())      # False: no peer access / IPC on this box, all_reduce only

    def step(how, ev=None):
        leg.step(partials.data_ptr())
        if ev is not None:
            ev[0].record(leg.stream)
        if how == 'peer':
            leg.eng.allreduce_chrom_dev(partials.data_ptr(), partials.numel())     # one kernel over NVLink peer memory
        else:
            dist.all_reduce(partials)
        if ev is not None:
            ev[1].record(leg.stream)
        leg.eng.finish_segments_dev(partials.data_ptr(), stats.data_ptr())

    res = {}
    with torch.cuda.stream(leg.stream):
        for how in (('peer', 'nccl') if peer_ok else ('nccl',)):
            for _ in range(warm):
                step(how)
            leg.eng.synchronize()
            barrier()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(leg.stream)
            for k in range(steps):
                step(how, evs[k])
            e1.record(leg.stream)
            leg.eng.synchronize()
            barrier()
            res[how] = (max_over_ranks(e0.elapsed_time(e1)), max_over_ranks(float(np.median([a.elapsed_time(b) for a, b in evs]))),
                        stats.cpu().numpy().copy())
    total_ms, ar_ms, st = res['peer' if peer_ok else 'nccl']
    finite = bool(np.isfinite(st).all())
    windows_here = hi - lo
    leg.close()
    torch.cuda.empty_cache()
    return {'workload': 'ONE configs[1] embryo, windows dealt to %d ranks by cost, chromosome partial sums summed across the GPUs every step' % world,
            'scaling': 'strong', 'value': len(reads) * subs * steps / (total_ms * 1e-3), 'unit': 'window-likelihoods/s',
            'ms_per_step': total_ms / steps, 'windows_rank0': int(windows_here), 'allreduce_bytes': int(partials.numel() * 8),
            'allreduce_ms_median': ar_ms, 'allreduce_share_of_step': ar_ms / (total_ms / steps),
            'collective': 'ldpgta_allreduce_chrom_dev: one kernel over NVLink peer memory (push to every peer, flag, poll locally, rank-order sum from own HBM)' if peer_ok
                          else 'torch.distributed all_reduce (NCCL): peer memory could not be mapped on this box',
            'nccl_comparison': {'ms_per_step': res['nccl'][0] / steps, 'allreduce_ms_median': res['nccl'][1],
                                'collective': 'torch.distributed all_reduce (NCCL) on the same device-resident partials'},
            'statistics_finite': finite}


def batch96_leg(L, torch, dist, device, rank, world, leg, barrier, max_over_ranks, segments):
    """ configs[4]: 96 embryos x 22 autosomes x 1000 subsamples per window, embryos dealt round-robin to the ranks,
    draws made on the device, only the chromosome statistics kept; the ranks' statistics are gathered at the end of the
    step.  Every embryo of a rank reuses the rank's synthetic read masks with its own seed (the work per embryo is that
    of a distinct embryo of the same shape; 96 distinct mask sets would need 265 GB). """
    n_embryos, subs = 96, 1000
    mine = list(range(rank, n_embryos, world))
    eng, n_seg = leg.eng, len(segments) - 1
    doff = np.arange(0, (leg.n_windows + 1) * subs, subs, dtype=np.int64)
    eng.set_windows(leg.woff, None, doff, leg.n, segment_offsets=segments)
    per_rank = (n_embryos + world - 1) // world
    stats = torch.zeros((per_rank, n_seg, 5, 3), dtype=torch.float64, device=device)
    gathered = torch.zeros((world, per_rank, n_seg, 5, 3), dtype=torch.float64, device=device) if dist is not None else None

    def step():
        for k, e in enumerate(mine):
            eng.bootstrap_stats_dev(leg.kind, 1000 + e, proportions=leg.proportions)
            eng.finish_segments_dev(None, stats[k].data_ptr())
        if dist is not None:
            dist.all_gather_into_tensor(gathered, stats)

    with torch.cuda.stream(leg.stream):
        step()                                                     # warm-up pass
        eng.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(leg.stream)
        step()
        e1.record(leg.stream)
        eng.synchronize()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    total = n_embryos * leg.n_windows * subs
    ok = bool(torch.isfinite(stats[:len(mine)]).all().item())
    return {'workload': 'configs[4]: 96 embryos x 22 autosomes x 1000 subsamples per window (4 reads per draw), embryos dealt to %d rank(s), '
                        'device draws, chromosome statistics only' % world,
            'value': total / (ms * 1e-3), 'unit': 'window-likelihoods/s', 'window_likelihoods': int(total), 'seconds': ms * 1e-3,
            'embryos_rank0': len(mine), 'statistics_finite': ok,
            'note': 'embryos of a rank share its synthetic read masks (distinct seeds); 96 distinct mask sets would need 265 GB'}


if __name__ == '__main__':
    sys.exit(main())
