"""
llr_bins -- the consumers of the per-window LLRs that the reference keeps in its plotting scripts
(SURVEY 8f rank 4): PLOT_PANEL.py:20-48, 50-74, 96-144 (chr_length, mean_and_std_of_mean_of_rnd_var, LLR,
load_likelihoods, bin_genomic_windows, binning), :146-245 (detect_crossovers, detect_crossovers_v2) and the
identical binning of BALANCED_ROC_CURVE.py:106-154.  Same names, arguments and return structures, so those
scripts can `from llr_bins import binning, detect_crossovers_v2` and keep their matplotlib code.

The reductions (LLRs of the five hypothesis pairs, per-bin mean and standard error) run on the device
(`ldpgta_window_stats` + `ldpgta_bin_stats`); `binning_pairs` goes from an LLR file's likelihoods to the binned
curves of all pairs without the per-draw LLRs returning to the host.  Bin bookkeeping and the crossover scan are
short sequential host loops over windows.  Plotting itself is out of scope.
"""
import bz2, gzip, pickle
from math import log

import numpy as np

from .engine import Engine, LLR_PAIRS

_HG38 = (248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
         135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
         46709983, 50818468, 156040895, 57227415)
_LENGTHS = {'chr' + str(c): n for c, n in zip([*range(1, 23), 'X', 'Y'], _HG38)}
_engines = {}


def _engine(device=0):
    if device not in _engines:
        _engines[device] = Engine([1], device=device)
    return _engines[device]


def chr_length(chr_id):
    """ hg38 chromosome length (PLOT_PANEL.py:20-28); KeyError for an unknown chromosome. """
    return _LENGTHS[chr_id]


def LLR(y, x):
    """ log(y/x) with the reference's sentinels (PLOT_PANEL.py:50-62). """
    if x and y:
        return log(y / x)
    if x:
        return -1.23456789
    if y:
        return +1.23456789
    return 0


def load_likelihoods(filename):
    """ (likelihoods, info) of an *.LLR.p[.gz|.bz2] file (PLOT_PANEL.py:64-74). """
    opener = {'bz2': bz2.open, 'gz': gzip.open, 'gzip': gzip.open}.get(filename.rpartition('.')[-1], open)
    with opener(filename, 'rb') as f:
        return pickle.load(f), pickle.load(f)


def _rows(A):
    rows = [tuple(v) for v in (A.values() if isinstance(A, dict) else A)]
    off = np.zeros(len(rows) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in rows])
    return np.fromiter((x for r in rows for x in r), dtype=np.float64, count=int(off[-1])), off


def mean_and_std_of_mean_of_rnd_var(A, device=0):
    """ PLOT_PANEL.py:37-48: rows of A are windows, columns bootstrap draws. """
    x, off = _rows(A)
    m, s = _engine(device).bin_stats(x, off, [[0, len(off) - 1]])[0, 0]
    return float(m), float(s)


def bin_genomic_windows(windows, chr_id, num_of_bins):
    """ {(i/num_of_bins, (i+1)/num_of_bins): (first window, end window) or None} -- PLOT_PANEL.py:96-119.
    A window belongs to the bin that holds its midpoint.  As in the reference, runs are half-open index ranges,
    the last window of the chromosome closes the final run without entering it, and bins without windows map to None. """
    width = chr_length(chr_id) / num_of_bins
    centre = [(a + b) / 2 for a, b in windows]
    label = [(i / num_of_bins, (i + 1) / num_of_bins) for i in range(num_of_bins)]
    table = {}
    cur = 0
    while cur < num_of_bins and not centre[0] < (cur + 1) * width:
        table[label[cur]] = None
        cur += 1
    cur = min(cur, num_of_bins - 1)
    first = 0
    for k, c in enumerate(centre):
        if width * cur <= c < width * (cur + 1):
            continue
        table[label[cur]] = (first, k)
        first = k
        while cur + 1 < num_of_bins:
            cur += 1
            if c < (cur + 1) * width:
                break
            table[label[cur]] = None
    last = len(centre) - 1
    for i in range(cur, num_of_bins):
        table[label[i]] = (first, last) if first != last else None
        first = last
    return table


def binning(LLRs_per_window, info, num_of_bins, device=0):
    """ PLOT_PANEL.binning (:121-144): (bin borders X, mean LLR Y, standard error E), None for empty bins.
    LLRs_per_window: {window: LLRs of its bootstrap draws} for one hypothesis pair. """
    bins = bin_genomic_windows([*LLRs_per_window], info['chr_id'], num_of_bins)
    x, off = _rows(LLRs_per_window)
    filled = [run for run in bins.values() if run]
    stats = iter(_engine(device).bin_stats(x, off, filled)[0].tolist()) if filled else iter(())
    Y, E = [], []
    for run in bins.values():
        m, s = next(stats) if run else (None, None)
        Y.append(m)
        E.append(s)
    return [*bins], Y, E


def binning_pairs(likelihoods, info, num_of_bins, pairs=LLR_PAIRS, device=0):
    """ {pair: (X, Y, E)} for hypothesis pairs straight from the likelihoods of an LLR file: the LLRs
    (PLOT_PANEL.py:304, 424) are formed on the device and stay there for the per-bin reduction. """
    unknown = [p for p in pairs if tuple(p) not in LLR_PAIRS]
    if unknown:
        raise KeyError('unsupported hypothesis pair(s): %r' % (unknown,))
    bins = bin_genomic_windows([*likelihoods], info['chr_id'], num_of_bins)
    counts = [len(v) for v in likelihoods.values()]
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    off[1:] = np.cumsum(counts)
    lik = np.array([tuple(l) for v in likelihoods.values() for l in v], dtype=np.float64).reshape(-1, 4)
    eng = _engine(device)
    eng.window_stats(lik, off, 0, want_llr=False)
    filled = [run for run in bins.values() if run]
    stats = eng.bin_stats(None, off, filled) if filled else np.empty((len(LLR_PAIRS), 0, 2))
    out = {}
    for pair in pairs:
        it = iter(stats[LLR_PAIRS.index(tuple(pair))].tolist())
        Y, E = [], []
        for run in bins.values():
            m, s = next(it) if run else (None, None)
            Y.append(m)
            E.append(s)
        out[tuple(pair)] = ([*bins], Y, E)
    return out


def _crossover_scan(genomic_windows, mean_of_LLRs, variance_of_LLRs, z_score, lookahead, recover):
    pos = [0.5 * (a + b) for a, b in genomic_windows]
    cum_m = np.cumsum(np.asarray(mean_of_LLRs, dtype=np.float64)).tolist()       # running sums, left to right
    cum_v = np.cumsum(np.asarray(variance_of_LLRs, dtype=np.float64)).tolist()
    n = min(len(pos), len(cum_m), len(cum_v))
    crossovers = {}
    peak = trough = None          # index of the running maximum / minimum since the last crossover
    prev, prev_sign = 0, 0

    def between(sign, i0, i1):
        """ the extremum of opposite kind that two consecutive same-kind crossovers must enclose (v2, :199-209) """
        best = None
        for i in range(*slice(i0 + 5, i1 - 4).indices(n)):        # Python slice semantics of triple[last+5:ind-4]
            if cum_v[i0] != cum_v[i] != cum_v[i1]:
                z = (cum_m[i] - cum_m[i0]) / (cum_v[i] - cum_v[i0]) ** .5 - (cum_m[i1] - cum_m[i]) / (cum_v[i1] - cum_v[i]) ** .5
                if best is None or (z < best[0] if sign > 0 else z > best[0]):
                    best = (z, i)
        if best is None:
            raise ValueError('%s() arg is an empty sequence' % ('min' if sign > 0 else 'max'))
        z, i = best
        kappa = min(abs(cum_m[i1] - cum_m[i]) / (cum_v[i1] - cum_v[i]) ** .5, abs(cum_m[i0] - cum_m[i]) / (cum_v[i] - cum_v[i0]) ** .5)
        print('Recovering skipped %s point:' % ('min' if sign > 0 else 'max'), (z, pos[i], cum_m[i], cum_v[i]))
        return pos[i], kappa

    for i in range(n):
        if peak is None or cum_m[i] > cum_m[peak]:
            peak = i
        if trough is None or cum_m[i] < cum_m[trough]:
            trough = i
        for sign, e in ((+1, peak), (-1, trough)):
            if e is None or i - e < lookahead:
                continue
            drop = sign * (cum_m[e] - cum_m[i])
            if not 0 < drop - z_score * (cum_v[i] - cum_v[e]) ** .5:
                continue
            for j in range(max(e - lookahead, prev), prev, -1):
                rise = sign * (cum_m[e] - cum_m[j])
                if 0 < rise - z_score * (cum_v[e] - cum_v[j]) ** .5:
                    if recover and prev_sign == sign:
                        x, kappa = between(sign, prev, e)
                        crossovers[x] = kappa
                    crossovers[pos[e]] = min(rise / (cum_v[e] - cum_v[j]) ** .5, drop / (cum_v[i] - cum_v[e]) ** .5)
                    peak = trough = None
                    prev, prev_sign = e, sign
                    break
    return crossovers


def detect_crossovers(genomic_windows, mean_of_LLRs, variance_of_LLRs, z_score=1.96, lookahead=20):
    """ {position: kappa} of the transitions between BPH and SPH regions (PLOT_PANEL.py:146-184). """
    return _crossover_scan(genomic_windows, mean_of_LLRs, variance_of_LLRs, z_score, lookahead, False)


def detect_crossovers_v2(genomic_windows, mean_of_LLRs, variance_of_LLRs, z_score=1.96, lookahead=20):
    """ detect_crossovers that also recovers a skipped extremum between two crossovers of the same kind
    (PLOT_PANEL.py:186-245). """
    return _crossover_scan(genomic_windows, mean_of_LLRs, variance_of_LLRs, z_score, lookahead, True)
