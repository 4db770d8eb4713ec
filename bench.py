#!/usr/bin/env python3
"""
bench.py -- bootstrapped window-likelihoods per second (BASELINE.json metric).

Workload (BASELINE.json configs[1]): one synthetic embryo, 22 autosomes at 0.1x, 100 kbp
windows (about 28.7k), ~150 informative reads per window, 5008-haplotype LD-structured panel,
32 bootstrap draws of 4 reads per window (the CLI defaults -w 100000 -s 32 -m 6 -M 4), all four
hypotheses (monosomy, disomy, SPH, BPH) per draw.  A "step" evaluates every (window, draw) of
the embryo once.  With N GPUs every rank owns one such embryo (weak scaling, no data-path
collective: (embryo, chromosome, window, draw) units are independent).

  value      draws/s with read masks, draw indices and outputs resident in HBM (CUDA events on the
             engine's stream, max over ranks)
  e2e        the same step through the public host API (Engine.likelihoods_uniform -> C-ABI
             ldpgta_likelihoods_uniform) from pinned host buffers: index upload + validation + kernel +
             result download inside the timed region; e2e.offsets_api_ms_per_step is the same step
             through ldpgta_likelihoods_batch (ragged API: also reads one offset per draw)
  roofline   dominant kernel k_small<4,8,10>: algorithmic bytes / measured launch time vs the
             measured HBM copy peak; int_roofline: algorithmic AND+POPC word-ops vs the POPC peak
  cpu_baseline  the CPython oracle port of the reference loop (ANEUPLOIDY_TEST.py:277-286) on a
             bounded sample of the same windows, one host core (the reference is single-threaded)

--impl reference times the oracle port on all host cores (one process per core) on bounded
samples of the same workload.
"""
import argparse
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
    chunk = 32
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
    """ For every window: effN = min(C(R, max_reads), subsamples) draws (ANEUPLOIDY_TEST.py:235-248) of
    n = min(R-1, max_reads) distinct reads (:231), as global read indices.  Uniform n only. """
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


def oracle_models(masks_rows, first_read, n_hap):
    """ The oracle's model object for one window (built once, outside the timed loop, like
    homogeneous.__init__ in the reference).  Reads are single-allele haplotypes (read_index, 'A'). """
    import ldpgta_oracle as O
    words = (n_hap + 63) // 64
    ints = [int.from_bytes(r[:words].tobytes(), 'little') for r in masks_rows]
    leg = [('chrS', first_read + i, 'C', 'A') for i in range(len(ints))]
    obs = [(first_read + i, 'read%d' % (first_read + i), 'A') for i in range(len(ints))]
    model = O.Homogeneous(obs, leg, {'ALL': ints}, {'ALL': n_hap}, {})
    reads = {first_read + i: [(first_read + i, 'A')] for i in range(len(ints))}
    return model, reads


def oracle_loop(model, reads, draws):
    """ The timed unit: the bootstrap inner loop (ANEUPLOIDY_TEST.py:282-286 minus the progress bar). """
    t0 = time.perf_counter()
    out = [tuple(model.get_likelihoods(tuple(reads[int(r)] for r in d))) for d in draws]
    return time.perf_counter() - t0, out


_W = {}


def _worker_init(masks, starts, idx, win, n_hap):
    _W.update(masks=masks, starts=starts, idx=idx, win=win, n_hap=n_hap, models={})


def _worker_step(windows):
    n = 0
    for w in windows:
        if w not in _W['models']:
            r0, r1 = int(_W['starts'][w]), int(_W['starts'][w + 1])
            _W['models'][w] = oracle_models(_W['masks'][r0:r1], r0, _W['n_hap']) + (_W['idx'][_W['win'] == w],)
        model, reads, draws = _W['models'][w]
        oracle_loop(model, reads, draws)
        n += len(draws)
    return n


def run_reference(args):
    """ --impl reference: the oracle port of the reference's CPU loop on all host cores (one process per
    core, windows dealt round-robin; model objects are built once per window outside the timed steps). """
    import multiprocessing as mp
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    wins, reads = workload_shape(args.depth, args.seed, min_reads=args.max_reads + 2, max_windows=args.windows)
    per_step_windows = min(max(cores * 8, 64), len(reads))   # bounded sample, the same windows every step
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
            'cpu_baseline': {'value': value, 'unit': 'window-likelihoods/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                             'python': sys.version.split()[0], 'popcount': 'int.bit_count (reference: gmpy2.popcount or byte LUT)'},
            'e2e': {'value': value, 'unit': 'window-likelihoods/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)
    return 0


def kernel_name(n, model):
    """ the kernel ldpgta_api.cu::run_uniform dispatches draws of n reads to """
    if n <= 4:
        return 'k_small<%d,8,10>' % n
    if n == 5 and model == 'homogeneous':
        return 'k_mid<5,5>'
    if n <= {'homogeneous': 12, 'recent': 12, 'distant': 10}[model] and not (model == 'distant' and n == 5):
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

    # ---- workload: one embryo per rank ------------------------------------------------------
    wins, reads = workload_shape(args.depth, args.seed + 1000 * rank, min_reads=args.max_reads + 2, max_windows=args.windows)
    n_windows, n_reads = len(reads), int(reads.sum())
    idx_host, win_of_draw = plan_draws(reads, args.subsamples, args.max_reads, args.seed + 1 + 1000 * rank)
    n_draws, n = idx_host.shape
    proportions = None
    if args.model == 'recent':                       # exploration: configs[3], two populations of 2504 haplotypes
        group_sizes, kind = [N_HAP // 2, N_HAP // 2], L.RECENT_ADMIXTURE
    elif args.model == 'distant':                    # exploration: the same panel with ancestral proportions 0.6 / 0.4
        group_sizes, kind, proportions = [N_HAP // 2, N_HAP // 2], L.DISTANT_ADMIXTURE, [0.6, 0.4]
    else:
        group_sizes, kind = [N_HAP], L.HOMOGENEOUS
    eng = L.Engine(group_sizes, device=local)
    all_masks = []
    for g, nh in enumerate(group_sizes):
        m = synth_masks(torch, device, reads, nh, args.seed + 1000 * rank + 77 * g)
        assert m.shape[1] == eng.padded_words[g]
        eng.attach_read_masks_dev(g, m.data_ptr(), n_reads, keepalive=m)
        all_masks.append(m)
    masks = all_masks[0]

    d_idx = torch.from_numpy(idx_host).to(device)
    d_out = torch.empty((n_draws, 4), dtype=torch.float64, device=device)
    stream = torch.cuda.ExternalStream(eng.stream, device=device)

    def step_resident():
        eng.likelihoods_uniform_dev(kind, n, d_idx.data_ptr(), n_draws, d_out.data_ptr(), proportions=proportions)

    # pinned host buffers for the end-to-end leg
    pin_idx = L.PinnedArray((n_draws, n), np.int32)
    pin_idx.array[:] = idx_host
    pin_out = L.PinnedArray((n_draws, 4), np.float64)
    offsets = np.arange(0, n_draws * n + 1, n, dtype=np.int64)

    def step_e2e():
        eng.likelihoods_uniform(kind, pin_idx.array, proportions=proportions, out=pin_out.array)

    def step_e2e_offsets():
        eng.likelihoods_batch(kind, pin_idx.array.reshape(-1), offsets, proportions=proportions, out=pin_out.array)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: resident timing -------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.5:               # untimed: let the SM clock leave idle before warm-up
        step_resident()
        eng.synchronize()
    for _ in range(args.warmup):
        step_resident()
    eng.synchronize()
    barrier()
    launches0 = eng.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for k in range(args.steps):
            step_resident()
            ev[k + 1].record(stream)
    eng.synchronize()
    barrier()
    launches = eng.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n_draws * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers through the public API --------------------------------------------
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n_draws * args.steps / e2e_s
    clocks = sampler.stop()
    first_pass = pin_out.array.copy()
    pin_out.array[:] = 0
    for _ in range(args.warmup):
        step_e2e_offsets()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_offsets()
    e2e_offsets_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    assert np.array_equal(first_pass, pin_out.array), 'the two host APIs disagree'
    resident = d_out.cpu().numpy()
    assert np.array_equal(resident, pin_out.array), 'resident and host-API results differ'

    # ---- multi-GPU: the only exchange of the path (chromosome LLR partial sums), outside the timed step ----
    stats_ms = None
    if dist is not None:
        off = np.arange(0, n_draws + 1, args.subsamples, dtype=np.int64)
        _, _, partials = eng.window_stats(pin_out.array, off, args.subsamples)
        tp = torch.from_numpy(partials).to(device)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.all_reduce(tp)
        torch.cuda.synchronize()
        stats_ms = 1e3 * (time.perf_counter() - t0)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- rooflines of the dominant kernel ------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    words = sum((nh + 63) // 64 for nh in group_sizes)
    bytes_per_draw = n * words * 8 + n * 4 + 32               # n*W*8 mask bytes + indices in, 32 B out
    wordops_per_draw = ((1 << n) - 1) * words
    k_ms = float(np.mean(kernel_ms))
    achieved_gbs = bytes_per_draw * n_draws / (k_ms * 1e-3) / 1e9
    int_peak = 2.3e12
    int_src = 'assumed 148 SM x 16 POPC.32/clk x 1.965 GHz / 2'
    try:
        ip = json.load(open(os.path.join(ROOT, 'profiles', 'INT_PEAKS.json')))
        int_peak = float(ip['wordop64_and_popc_add']['ops_per_s'])
        int_src = 'measured (profiles/INT_PEAKS.json, tools/int_peaks.cu: AND+POPC+ADD on 64-bit words)'
    except Exception:
        pass
    achieved_wordops = wordops_per_draw * n_draws / (k_ms * 1e-3)
    traffic = None                                            # DRAM bytes per launch from the committed ncu capture of this kernel
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'))).get('k_small<%d,8,10>' % n)
        if t and n <= 4:
            traffic = t['dram_bytes_per_launch'] * n_draws / t['draws_per_launch']
    except Exception:
        pass

    # ---- CPU baseline: oracle port on a bounded sample, one core; doubles as the parity check -------
    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline and world == 1 and args.model == 'homogeneous':     # rank 0 at N=1 only
        spent, n_eval, w = 0.0, 0, 0
        ref_rows, got_rows = [], []
        starts = np.concatenate([[0], np.cumsum(reads)])
        while spent < args.cpu_seconds and w < n_windows:
            r0, r1 = int(starts[w]), int(starts[w + 1])
            model, rd = oracle_models(masks[r0:r1].cpu().numpy().view(np.uint64), r0, N_HAP)
            sel = slice(w * args.subsamples, (w + 1) * args.subsamples)          # draws are grouped by window
            assert (win_of_draw[sel] == w).all()
            t, vals = oracle_loop(model, rd, idx_host[sel])
            ref_rows.append(np.array(vals, dtype=np.float64).reshape(-1, 4))
            got_rows.append(resident[sel])
            spent += t; n_eval += args.subsamples; w += 1
        sample_windows = list(range(w))
        ref_v, got_v = np.concatenate(ref_rows), np.concatenate(got_rows)
        parity = float(np.max(np.abs(ref_v - got_v) / np.maximum(np.abs(ref_v), 1e-300)))
        cpu_baseline = {'value': n_eval / spent, 'unit': 'window-likelihoods/s', 'cores': 1, 'kind': 'port',
                        'sample': 'first %d windows x %d draws (%d draws, %.1f s) of the same embryo' % (len(sample_windows), args.subsamples, n_eval, spent),
                        'python': sys.version.split()[0], 'popcount': 'int.bit_count (reference: gmpy2.popcount or byte LUT)',
                        'host_cores_available': os.cpu_count()}

    line = {
        'metric': 'bootstrapped window-likelihoods/sec', 'value': value, 'unit': 'window-likelihoods/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u64 and/popcount + f64 polynomials', 'data': 'synthetic',
        'config': workload_config(args, n_windows, n_reads),
        'draws_per_step_per_gpu': int(n_draws),
        'e2e': {'value': e2e_value, 'unit': 'window-likelihoods/s', 'h2d_bytes_per_step': int(pin_idx.array.nbytes),
                'd2h_bytes_per_step': int(pin_out.array.nbytes), 'ms_per_step': 1e3 * e2e_s / args.steps,
                'api': 'Engine.likelihoods_uniform -> ldpgta_likelihoods_uniform (pinned host buffers)',
                'offsets_api_ms_per_step': e2e_offsets_ms},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': kernel_name(n, args.model), 'achieved': achieved_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': achieved_gbs / hbm_peak, 'traffic': traffic, 'peak_source': peak_src,
                     'algorithmic_bytes_per_draw': bytes_per_draw, 'kernel_ms': k_ms},
        'int_roofline': {'bound': 'popc', 'achieved': achieved_wordops, 'peak': int_peak, 'unit': 'word AND+POPC/s',
                         'frac': achieved_wordops / int_peak, 'peak_source': int_src, 'wordops_per_draw': wordops_per_draw},
        'cpu_baseline': cpu_baseline,
        'parity_max_rel_err_vs_oracle': parity,
        'clocks': clocks,
    }
    if stats_ms is not None:
        line['chromosome_partials_allreduce_ms'] = stats_ms
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
