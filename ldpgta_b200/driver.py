"""
Drop-in for the driver ANEUPLOIDY_TEST.py: same functions, arguments, return
structures, pickle formats and CLI.  The bootstrap loop (:277-286) is executed
as ONE batched device call: every draw of every window is sampled first, with
the same `random.sample` call sequence as the reference (so a given seed and
PYTHONHASHSEED reproduce the reference's draws), then all draws are evaluated
by libldpgta.so; LLRs and their per-window / per-chromosome reductions
(statistics(), :289-325) also run on the device.
"""
import bz2
import collections
import gzip
import os
import pickle
import random
import re
import time
from itertools import product
from math import comb
from statistics import StatisticsError, mean, pstdev, variance

import numpy as np

from . import _native as N
from .engine import Engine, LLR_PAIRS, finish_chromosome
from .models import ImplicitModels, distant_admixture, homogeneous, likelihoods_tuple, recent_admixture

leg_tuple = collections.namedtuple('leg_tuple', ('chr_id', 'pos', 'ref', 'alt'))
obs_tuple = collections.namedtuple('obs_tuple', ('pos', 'read_id', 'base'))
comb_tuple = collections.namedtuple('comb_tuple', ('ref', 'alt', 'hap'))


def popcount(x):
    return x.bit_count()


def mean_and_var(x):
    """ Mean and sample variance (ANEUPLOIDY_TEST.py:51-56). """
    cache = tuple(x)
    m = mean(cache)
    return m, variance(cache, xbar=m)


def mean_and_std(x):
    """ Mean and population standard deviation (ANEUPLOIDY_TEST.py:58-63). """
    cache = tuple(x)
    m = mean(cache)
    return m, pstdev(cache, mu=m)


def mean_and_std_of_mean_of_rnd_var(A):
    """ ANEUPLOIDY_TEST.py:65-76 on host containers (the device version is Engine.window_stats). """
    if type(A) == dict:
        A = tuple(tuple(i) for i in A.values())
    M, Nn = len(A), len(A[0])
    mu = sum(sum(row) / Nn for row in A)
    std = (sum((sum(col) - mu) ** 2 for col in zip(*A)) / (Nn - 1)) ** .5 / M
    return mu / M, std


def LLR(y, x):
    """ log(y/x) with the sentinels of ANEUPLOIDY_TEST.py:78-90. """
    from math import log
    if x and y:
        return log(y / x)
    if x and not y:
        return -1.23456789
    if not x and y:
        return +1.23456789
    return 0


def invert(x, n):
    """ Inverts the n low bits of a non-negative integer (ANEUPLOIDY_TEST.py:92-94). """
    return x ^ ((1 << n) - 1)


class supporting_dictionaries:
    """ reads / overlaps / positions / per-group panel rows / read scores, as
    ANEUPLOIDY_TEST.supporting_dictionaries (:96-189). Host-side in this round. """

    def __init__(self, obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, ancestral_makeup, min_HF, examine=None):
        """ `examine` (a model object of models.py holding the device-resident allele rows) moves the read
        scores to the GPU; without it they are computed on the host exactly as the reference does. """
        self.number_of_haplotypes_per_group = number_of_haplotypes_per_group
        self.min_HF = min_HF
        if type(ancestral_makeup) == dict:
            self.ancestral_makeup = {group2: ancestral_makeup[group2] for group2 in hap_tab_per_group}
        else:
            self.ancestral_makeup = {group2: 1 / len(ancestral_makeup) for group2 in hap_tab_per_group}
        self._leg_tab = leg_tab
        wanted = {row[0] for row in obs_tab}           # leg_dict restricted to observed positions; the full one is built on demand
        observed_leg = {l[1]: (l[2], l[3]) for l in leg_tab if l[1] in wanted}
        self.reads = self.build_reads_dict(obs_tab, observed_leg)
        self.overlaps = self.build_overlaps_dict(obs_tab, observed_leg)
        self.positions = (*self.overlaps.keys(),)
        if examine is not None:
            self.hap_tab_per_group = None              # the panel rows live on the device
            self.score = self.build_score_dict_device(self.reads, examine, self.number_of_haplotypes_per_group,
                                                      self.ancestral_makeup, self.min_HF)
        else:
            self.hap_tab_per_group = self.build_hap_tab_per_group(hap_tab_per_group, leg_tab, self.positions)
            self.score = self.build_score_dict(self.reads, self.hap_tab_per_group, self.number_of_haplotypes_per_group,
                                               self.ancestral_makeup, self.min_HF)

    @property
    def leg_dict(self):
        """ position -> (ref, alt) for the whole legend (ANEUPLOIDY_TEST.py:110). """
        if not hasattr(self, '_leg_dict'):
            self._leg_dict = {l[1]: (l[2], l[3]) for l in self._leg_tab}
        return self._leg_dict

    def build_reads_dict(self, obs_tab, leg_dict):
        reads = collections.defaultdict(list)
        for pos, read_id, base in obs_tab:
            if base in leg_dict.get(pos, ()):
                reads[read_id].append((pos, base))
        return reads

    def build_overlaps_dict(self, obs_tab, leg_dict):
        overlaps = collections.defaultdict(list)
        for pos, read_id, base in obs_tab:
            if base in leg_dict.get(pos, ()):
                overlaps[pos].append(read_id)
        return overlaps

    def build_hap_tab_per_group(self, hap_tab_per_group, leg_tab, positions):
        wanted = set(positions)
        return {group2: {leg[1]: hap for leg, hap in zip(leg_tab, hap_tab) if leg[1] in wanted}
                for group2, hap_tab in hap_tab_per_group.items()}

    def build_score_dict_device(self, reads_dict, examine, number_of_haplotypes_per_group, ancestral_makeup, min_HF):
        """ build_score_dict (ANEUPLOIDY_TEST.py:151-189) on the GPU: one warp per read enumerates the 2^k
        allele/complement combinations of its common SNPs (ldpgta_read_scores). """
        alpha = {g: ancestral_makeup[g] / n for g, n in number_of_haplotypes_per_group.items()}
        weight = list(alpha.values())[-1]          # the reference weights every group by the last alpha (:181-184)
        index = examine._allele_index
        allele_idx, offsets, ids = [], [0], []
        for read_id, alleles in reads_dict.items():
            allele_idx.extend(index[a] for a in alleles)
            offsets.append(len(allele_idx))
            ids.append(read_id)
        scores = examine._engine.read_scores(allele_idx, offsets, [alpha[g] for g in examine._groups], weight, min_HF)
        return dict(zip(ids, scores.tolist()))

    def build_score_dict(self, reads_dict, hap_tab_per_group, number_of_haplotypes_per_group, ancestral_makeup, min_HF):
        """ Number of allele/complement combinations of a read whose effective haplotype
        frequency lies in [min_HF, 1-min_HF]; SNPs with effective allele frequency outside
        [0.01, 0.99] are ignored (ANEUPLOIDY_TEST.py:151-189). """
        max_HF = 1 - min_HF
        alpha = {g: ancestral_makeup[g] / n for g, n in number_of_haplotypes_per_group.items()}
        ones = {g: (1 << n) - 1 for g, n in number_of_haplotypes_per_group.items()}
        # :181-184 consumes per-group generators after the comprehension that made them has
        # finished, so every group ends up weighted by the LAST group's alpha; kept for parity.
        a_last = list(alpha.values())[-1]
        score = {}
        for read_id, alleles in reads_dict.items():
            rows = collections.defaultdict(list)
            for pos, _base in alleles:
                freq = sum(a * hap_tab_per_group[g][pos].bit_count() for g, a in alpha.items())
                if 0.01 <= freq <= 0.99:
                    for g, b in ones.items():
                        h = hap_tab_per_group[g][pos]
                        rows[g].append((h, h ^ b))
            if rows:
                per_group = []
                for g in alpha:
                    vals = []
                    for choice in product(*rows[g]):
                        m = choice[0]
                        for r in choice[1:]:
                            m &= r
                        vals.append(a_last * m.bit_count())
                    per_group.append(vals)
                score[read_id] = sum(min_HF <= sum(t) <= max_HF for t in zip(*per_group))
            else:
                score[read_id] = 0
        return score


def build_gw_dict(obs, window_size, offset, min_reads, max_reads, min_score):
    """ Genomic windows -> read IDs (ANEUPLOIDY_TEST.py:191-224), fixed or adaptive.  The window borders come from the
    array implementation packed.build_windows (one implementation, pinned to the reference's windows in
    tests/test_packed.py); this wrapper only turns them into the reference's dict.  A window's value is a tuple made
    from a set that received the qualifying reads position by position, as the reference fills it, so that for a given
    PYTHONHASHSEED the tuples come out in the reference's order (random.sample indexes into them, :231). """
    from .packed import build_windows
    names = list(obs.reads)
    row_of = {name: i for i, name in enumerate(names)}
    counts = np.fromiter((len(ids) for ids in obs.overlaps.values()), dtype=np.int64, count=len(obs.overlaps))
    obs_pos = np.repeat(np.fromiter(obs.overlaps.keys(), dtype=np.int64, count=len(obs.overlaps)), counts)
    obs_read = np.fromiter((row_of[r] for ids in obs.overlaps.values() for r in ids), dtype=np.int64, count=int(counts.sum()))
    scores = np.fromiter((obs.score[name] for name in names), dtype=np.int64, count=len(names))
    win = build_windows(obs_pos, obs_read, scores, window_size, offset, min_reads, min_score)
    keep = scores[obs_read] >= min_score
    lo = np.searchsorted(obs_pos, win.borders[:, 0], side='left').tolist()
    hi = np.searchsorted(obs_pos, win.borders[:, 1], side='right').tolist()
    gw_dict = {}
    for (a, b), i, j in zip(win.borders.tolist(), lo, hi):
        members = set()
        members.update(names[r] for r in obs_read[i:j][keep[i:j]].tolist())
        gw_dict[a, b] = tuple(members)
    return gw_dict


def pick_reads(reads_dict, read_IDs, max_reads):
    """ One draw without replacement (ANEUPLOIDY_TEST.py:226-233): max_reads reads, or all but one when the window
    holds no more than that; returns the reads themselves (lists of alleles). """
    size = min(max_reads, len(read_IDs) - 1)
    return tuple(map(reads_dict.__getitem__, random.sample(read_IDs, size)))


def effective_number_of_subsamples(num_of_reads, min_reads, max_reads, subsamples):
    """ ANEUPLOIDY_TEST.py:235-248. """
    if min_reads <= num_of_reads and num_of_reads > max_reads:
        return min(comb(num_of_reads, max_reads), subsamples)
    if min_reads <= num_of_reads <= max_reads:
        return min(num_of_reads, subsamples)
    return 0


def make_examiner(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, models_dict, device=0, panel=None):
    """ Model selection of ANEUPLOIDY_TEST.py:263-273. """
    if type(ancestral_makeup) == set and len(ancestral_makeup) == 1:
        print('Assuming one ancestral population: %s.' % next(iter(ancestral_makeup)))
        return homogeneous(obs_tab, leg_tab, hap_tab, number_of_haplotypes, models_dict, device=device, panel=panel)
    if type(ancestral_makeup) == set and len(ancestral_makeup) == 2:
        print('Assuming recent-admixture between %s and %s.' % tuple(ancestral_makeup))
        return recent_admixture(obs_tab, leg_tab, hap_tab, number_of_haplotypes, models_dict, device=device, panel=panel)
    if type(ancestral_makeup) == dict:
        print('Assuming the following ancestral makeup:', ", ".join("{:.1f}% {}".format(100 * v, k) for k, v in ancestral_makeup.items()))
        return distant_admixture(obs_tab, leg_tab, hap_tab, number_of_haplotypes, models_dict, ancestral_makeup, device=device, panel=panel)
    raise Exception('error: unsupported ancestral-makeup.')


def plan_draws(genomic_windows, min_reads, max_reads, subsamples, replay=None):
    """ Every draw of the bootstrap loop, in the reference's order: for each window with
    effN > 0, effN calls of random.sample(read_IDs, min(R-1, max_reads)) (:277-284).
    `replay` = {window: [tuple(read ids), ...]} substitutes recorded draws. """
    draws = {}
    for window, read_IDs in genomic_windows.items():
        effN = effective_number_of_subsamples(len(read_IDs), min_reads, max_reads, subsamples)
        if effN:
            if replay is not None:
                draws[window] = [tuple(d) for d in replay[window][:effN]]
            else:
                k = min(len(read_IDs) - 1, max_reads)
                draws[window] = [tuple(random.sample(read_IDs, k)) for _ in range(effN)]
    return draws


def evaluate_draws(examine, reads_dict, draws):
    """ All draws in one device batch.  Returns {window: tuple(likelihoods_tuple)}. """
    read_index = {}
    allele_idx, read_offsets = [], [0]
    rows_of = examine._rows_of
    flat_idx, draw_offsets, sizes = [], [0], []
    for window, ds in draws.items():
        for ids in ds:
            for r in ids:
                i = read_index.get(r)
                if i is None:
                    i = read_index[r] = len(read_offsets) - 1
                    allele_idx.extend(rows_of((reads_dict[r],))[0])
                    read_offsets.append(len(allele_idx))
                flat_idx.append(i)
            draw_offsets.append(len(flat_idx))
            sizes.append(len(ids))
    if not sizes:
        return {}
    for n in set(sizes):
        if n not in (2, 3, 4):
            examine._check_table(n)
        if n == 5 and examine.kind == N.DISTANT_ADMIXTURE:
            raise AttributeError("'distant_admixture' object has no attribute 'likelihoods5'")
    eng = examine._engine
    eng.build_read_masks(allele_idx, read_offsets)
    lik = eng.likelihoods_batch(examine.kind, flat_idx, draw_offsets, proportions=examine._proportions())
    T = examine.tuple_type
    out, k = {}, 0
    rows = lik.tolist()
    for window, ds in draws.items():
        out[window] = tuple(T(*rows[k + j]) for j in range(len(ds)))
        k += len(ds)
    return out


def plan_draw_indices(genomic_windows, min_reads, max_reads, subsamples, replay=None, fast_rng=None):
    """ The same draws as plan_draws, as positions inside each window's read tuple: {window: int array [effN, k]}.
    random.sample(range(R), k) consumes the global MT19937 stream exactly like random.sample(read_IDs, k) and
    returns the positions it would have picked, so the draws are the reference's for a given seed.
    `fast_rng` (a numpy Generator) switches to vectorised sampling without replacement for throughput runs
    (e.g. 1000 subsamples per window); those draws are statistically equivalent, not the reference's stream. """
    plan = {}
    for window, read_IDs in genomic_windows.items():
        R = len(read_IDs)
        effN = effective_number_of_subsamples(R, min_reads, max_reads, subsamples)
        if not effN:
            continue
        k = min(R - 1, max_reads)
        if replay is not None:
            pos = {r: i for i, r in enumerate(read_IDs)}
            plan[window] = np.array([[pos[r] for r in d] for d in replay[window][:effN]], dtype=np.int64).reshape(effN, -1)
        elif fast_rng is not None:
            plan[window] = np.argsort(fast_rng.random((effN, R)), axis=1)[:, :k]
        else:
            population = range(R)
            plan[window] = np.array([random.sample(population, k) for _ in range(effN)], dtype=np.int64).reshape(effN, k)
    return plan


def evaluate_draw_indices(examine, reads_dict, genomic_windows, plan, with_stats=False):
    """ evaluate_draws for index plans: one read-mask row per read of the evaluated windows, one device batch.
    with_stats=True returns (likelihoods, window_mean_var[5][n_windows][2], chromosome[5][3]) from ONE device-resident pass
    (ldpgta_set_windows + ldpgta_replay_stats): the likelihoods feed the LLR reductions without leaving HBM. """
    if not plan:
        return ({}, None, None) if with_stats else {}
    index = examine._allele_index
    read_row = {}
    allele_idx, read_offsets = [], [0]
    chunks, sizes = [], set()
    for window, local in plan.items():
        rows = np.empty(len(genomic_windows[window]), dtype=np.int64)
        for j, r in enumerate(genomic_windows[window]):
            i = read_row.get(r)
            if i is None:
                i = read_row[r] = len(read_offsets) - 1
                allele_idx.extend(index[a] for a in reads_dict[r])
                read_offsets.append(len(allele_idx))
            rows[j] = i
        chunks.append(rows[local])                       # [effN, k] global read rows
        sizes.add(local.shape[1])
    for n in sizes:
        if n not in (2, 3, 4):
            examine._check_table(n)
        if n == 5 and examine.kind == N.DISTANT_ADMIXTURE:
            raise AttributeError("'distant_admixture' object has no attribute 'likelihoods5'")
    counts = np.array([c.shape[0] for c in chunks], dtype=np.int64)
    widths = np.repeat(np.array([c.shape[1] for c in chunks], dtype=np.int64), counts)
    flat = np.concatenate([c.ravel() for c in chunks]).astype(np.int32)
    offsets = np.zeros(len(widths) + 1, dtype=np.int64)
    np.cumsum(widths, out=offsets[1:])
    eng = examine._engine
    eng.build_read_masks(allele_idx, read_offsets)
    mean_var = chrom = None
    if with_stats:
        draw_offsets = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=draw_offsets[1:])
        eng.set_windows(None, None, draw_offsets, np.array([c.shape[1] for c in chunks], dtype=np.int32))     # replay-only plan
        res = eng.replay_stats(examine.kind, flat, proportions=examine._proportions(), want=('lik', 'mean_var', 'stats'))
        lik, mean_var, chrom = res['lik'], res['mean_var'], res['stats'][0]
    else:
        lik = eng.likelihoods_batch(examine.kind, flat, offsets, proportions=examine._proportions())
    T = examine.tuple_type
    out, k = {}, 0
    rows = lik.tolist()
    for window, c in zip(plan, counts.tolist()):
        out[window] = tuple(T(*rows[k + j]) for j in range(c))
        k += c
    return (out, mean_var, chrom) if with_stats else out


def bootstrap(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, models_dict, window_size, subsamples,
              offset, min_reads, max_reads, min_score, min_HF, replay=None, device=0, panel=None, fast_rng=None, _stats_out=None):
    """ ANEUPLOIDY_TEST.bootstrap (:250-287): (likelihoods, genomic_windows, fraction_of_matches).
    `panel`: an engine.ResidentPanel of this chromosome shared by many samples (for a dict makeup: groups in the makeup's order);
    `fast_rng`: numpy Generator for vectorised draws (throughput runs; not the reference's random stream). """
    examine = make_examiner(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, models_dict, device=device, panel=panel)
    try:
        obs = supporting_dictionaries(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, min_HF, examine=examine)
        genomic_windows = build_gw_dict(obs, window_size, offset, min_reads, max_reads, min_score)
        plan = plan_draw_indices(genomic_windows, min_reads, max_reads, subsamples, replay=replay, fast_rng=fast_rng)
        if _stats_out is not None and plan and min(len(v) for v in plan.values()) >= 2:
            # aneuploidy_test: statistics() from the same device pass (filled into _stats_out for the caller)
            likelihoods, mean_var, chrom = evaluate_draw_indices(examine, obs.reads, genomic_windows, plan, with_stats=True)
            _stats_out.update(device_statistics(likelihoods, genomic_windows, mean_var, chrom))
        else:
            likelihoods = evaluate_draw_indices(examine, obs.reads, genomic_windows, plan)
        return likelihoods, genomic_windows, examine.fraction_of_matches
    finally:
        examine.close()


_stats_engine = {}


def _engine_for_stats(device=0):
    if device not in _stats_engine:
        _stats_engine[device] = Engine([1], device=device)
    return _stats_engine[device]


def chromosome_arrays(likelihoods):
    """ {window: tuple(likelihoods_tuple)} -> (float64[n_draws, 4], window offsets, min draws, first window's draws). """
    counts = [len(v) for v in likelihoods.values()]
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    off[1:] = np.cumsum(counts)
    lik = np.array([tuple(l) for v in likelihoods.values() for l in v], dtype=np.float64).reshape(-1, 4)
    return lik, off, min(counts), counts[0]


def device_statistics(likelihoods, genomic_windows, mean_var, chrom):
    """ The dictionary of ANEUPLOIDY_TEST.statistics (:289-325) from reductions the device has already made:
    mean_var[5][n_windows][2] in the order of `likelihoods`, chrom[5][3]. """
    window_size_mean, window_size_std = mean_and_std((j - i + 1 for (i, j) in likelihoods))
    reads_mean, reads_std = mean_and_std((len(read_IDs) for window, read_IDs in genomic_windows.items() if window in likelihoods))
    mv = np.asarray(mean_var).tolist()
    per_window = {pair: {w: tuple(mv[p][k]) for k, w in enumerate(likelihoods)} for p, pair in enumerate(LLR_PAIRS)}
    per_chrom = {pair: {'mean_of_mean': float(chrom[p][0]), 'std_of_mean': float(chrom[p][1]),
                        'fraction_of_negative_LLRs': float(chrom[p][2])} for p, pair in enumerate(LLR_PAIRS)}
    return {'num_of_windows': len(likelihoods), 'reads_mean': reads_mean, 'reads_std': reads_std,
            'window_size_mean': window_size_mean, 'window_size_std': window_size_std,
            'LLRs_per_genomic_window': per_window, 'LLRs_per_chromosome': per_chrom}


def statistics(likelihoods, genomic_windows, device=0):
    """ ANEUPLOIDY_TEST.statistics (:289-325) with the LLRs and their reductions on the device. """
    if not likelihoods:
        return {'num_of_windows': 0}
    window_size_mean, window_size_std = mean_and_std((j - i + 1 for (i, j) in likelihoods))
    reads_mean, reads_std = mean_and_std((len(read_IDs) for window, read_IDs in genomic_windows.items() if window in likelihoods))
    lik, off, n_cols, first = chromosome_arrays(likelihoods)
    if min(np.diff(off)) < 2:
        raise StatisticsError('variance requires at least two data points')     # statistics.variance, via :55
    _llr, mean_var, partials = _engine_for_stats(device).window_stats(lik, off, n_cols)
    chrom = finish_chromosome(partials, n_cols, first)
    windows = list(likelihoods)
    per_window = {pair: {w: (float(mean_var[p, k, 0]), float(mean_var[p, k, 1])) for k, w in enumerate(windows)}
                  for p, pair in enumerate(LLR_PAIRS)}
    per_chrom = {pair: {'mean_of_mean': float(chrom[p, 0]), 'std_of_mean': float(chrom[p, 1]),
                        'fraction_of_negative_LLRs': float(chrom[p, 2])} for p, pair in enumerate(LLR_PAIRS)}
    return {'num_of_windows': len(likelihoods), 'reads_mean': reads_mean, 'reads_std': reads_std,
            'window_size_mean': window_size_mean, 'window_size_std': window_size_std,
            'LLRs_per_genomic_window': per_window, 'LLRs_per_chromosome': per_chrom}


def packed_bootstrap(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, window_size, subsamples, offset,
                     min_reads, max_reads, min_score, min_HF, device=0, panel=None, rng=None):
    """ bootstrap() + statistics() through packed.PackedSample; returns the reference's structures
    (likelihoods, genomic_windows, fraction_of_matches, statistics dict). """
    from .engine import ResidentPanel
    from .packed import PackedSample
    own_panel = panel is None
    if own_panel:
        order = list(ancestral_makeup) if type(ancestral_makeup) == dict else list(hap_tab)
        panel = ResidentPanel(leg_tab, {g: hap_tab[g] for g in order}, number_of_haplotypes, device=device)
    try:
        with PackedSample(obs_tab, panel, ancestral_makeup, min_HF) as sample:
            res = sample.run(window_size, subsamples, offset, min_reads, max_reads, min_score, rng=rng)
            cls = {N.HOMOGENEOUS: homogeneous, N.RECENT_ADMIXTURE: recent_admixture, N.DISTANT_ADMIXTURE: distant_admixture}[sample.kind]
            likelihoods, genomic_windows, matches = res.as_reference(cls.tuple_type)
    finally:
        if own_panel:
            panel.close()
    if not likelihoods:
        return likelihoods, genomic_windows, matches, {'num_of_windows': 0}
    if min(len(v) for v in likelihoods.values()) < 2:
        raise StatisticsError('variance requires at least two data points')
    window_size_mean, window_size_std = mean_and_std((j - i + 1 for (i, j) in likelihoods))
    reads_mean, reads_std = mean_and_std((len(read_IDs) for window, read_IDs in genomic_windows.items() if window in likelihoods))
    mv = res.window_mean_var.tolist()
    per_window = {pair: {w: tuple(mv[p][k]) for k, w in enumerate(likelihoods)} for p, pair in enumerate(LLR_PAIRS)}
    stats = {'num_of_windows': len(likelihoods), 'reads_mean': reads_mean, 'reads_std': reads_std,
             'window_size_mean': window_size_mean, 'window_size_std': window_size_std,
             'LLRs_per_genomic_window': per_window, 'LLRs_per_chromosome': res.llr_per_chromosome()}
    return likelihoods, genomic_windows, matches, stats


def print_summary(obs_filename, info):
    """ The console report of ANEUPLOIDY_TEST.print_summary (:327-343), same text. """
    S = info['statistics']
    makeup = info['ancestral_makeup']
    if isinstance(makeup, dict):
        makeup_text = ', '.join('%.1f%% %s' % (100 * share, group) for group, share in makeup.items())
    else:
        makeup_text = ', '.join(makeup)
    matched_text = ', '.join('%s: %.1f%%' % (group, 100 * frac) for group, frac in S['matched_alleles'].items())
    report = ['', 'Filename: %s' % obs_filename, '', 'Summary statistics', '------------------',
              'Chromosome ID: %s, Depth: %.2f.' % (info['chr_id'], info['depth']),
              'Number of genomic windows: %d, Mean and standard error of genomic window size: %d, %d.'
              % tuple(S.get(k, 0) for k in ('num_of_windows', 'window_size_mean', 'window_size_std')),
              'Mean and standard error of meaningful reads per genomic window: %.1f, %.1f.' % (S.get('reads_mean', 0), S.get('reads_std', 0)),
              'Ancestral makeup: %s, Fraction of alleles matched to the reference panel: %s.' % (makeup_text, matched_text)]
    for (num, den), chrom in (S.get('LLRs_per_chromosome') or {}).items():
        report += ['--- Chromosome-wide LLR between %s and %s ----' % (num, den),
                   'Mean LLR: %.3f, Standard error of the mean LLR: %.3f' % (chrom['mean_of_mean'], chrom['std_of_mean']),
                   'Fraction of genomic windows with a negative LLR: %.3f' % chrom['fraction_of_negative_LLRs']]
    print('\n'.join(report))


def llr_filename(obs_filename, output_filename, compress):
    """ Output name rule of ANEUPLOIDY_TEST.save_results (:352-358): by default the observation file's base name with
    everything from its last `obs.p` on replaced by `LLR.p`; gz / bz2 add their extension to either form. """
    ext = '.' + compress if compress in ('bz2', 'gz') else ''
    if output_filename != '':
        return output_filename + ext
    base = obs_filename.rsplit('/', 1)[-1]
    stem = re.match('(.*)obs.p', base)
    return stem.group(1) + 'LLR.p' + ext if stem else base


def save_results(likelihoods, info, compress, obs_filename, output_filename, output_dir):
    """ Two consecutive protocol-4 pickles (likelihoods, info), optionally gz / bz2 (ANEUPLOIDY_TEST.py:345-366);
    returns the path written. """
    directory = output_dir.rstrip('/') + '/'
    if not os.path.exists(directory):
        os.makedirs(directory)
    path = directory + llr_filename(obs_filename, output_filename, compress)
    opener = bz2.open if compress == 'bz2' else gzip.open if compress == 'gz' else open
    with opener(path, 'wb') as f:
        for obj in (likelihoods, info):
            pickle.dump(obj, f, protocol=4)
    return path


def _opener(filename):
    return {'bz2': bz2.open, 'gz': gzip.open}.get(filename.rsplit('.', 1)[1], open)


def aneuploidy_test(obs_filename, leg_filename, hap_filename, ancestral_makeup, window_size, subsamples, offset, min_reads,
                    max_reads, min_score, min_HF, output_filename, compress, **kwargs):
    """ ANEUPLOIDY_TEST.aneuploidy_test (:368-429): same arguments, files and return value.
    kwargs: seed, model (MODELS*.p path; optional here -- the device needs no tables), output_dir, device,
    panel (ResidentPanel shared by samples), fast_draws (vectorised numpy draws instead of random.sample),
    packed (the array-native route of packed.py: same files and results, draws from a numpy Generator). """
    time0 = time.time()
    random.seed(a=kwargs.get('seed', None), version=2)

    with _opener(obs_filename)(obs_filename, 'rb') as obs_in:
        obs_tab = pickle.load(obs_in)
        info = pickle.load(obs_in)
    with _opener(hap_filename)(hap_filename, 'rb') as hap_in:
        hap_tab = pickle.load(hap_in)
        number_of_haplotypes = pickle.load(hap_in)
    with _opener(leg_filename)(leg_filename, 'rb') as leg_in:
        leg_tab = pickle.load(leg_in)

    models_filename = kwargs.get('model', None)
    if models_filename is not None:
        with _opener(models_filename)(models_filename, 'rb') as model_in:
            models_dict = pickle.load(model_in)
    else:
        models_dict = ImplicitModels()

    ancestry = set(ancestral_makeup.keys()) if type(ancestral_makeup) == dict else set(ancestral_makeup)
    assert ancestry == {*hap_tab}, 'error: the given ancestral makup does match the superpopulations (group2) in the reference panel.'

    device = kwargs.get('device', 0)
    fast_rng = np.random.default_rng(kwargs.get('seed', None)) if kwargs.get('fast_draws', False) else None
    if kwargs.get('packed', False):
        likelihoods, genomic_windows, matched_alleles, stats = packed_bootstrap(
            obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, window_size, subsamples, offset, min_reads,
            max_reads, min_score, min_HF, device=device, panel=kwargs.get('panel', None),
            rng=np.random.default_rng(kwargs.get('seed', None)))
    else:
        stats = {}
        likelihoods, genomic_windows, matched_alleles = bootstrap(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup,
                                                                  models_dict, window_size, subsamples, offset, min_reads,
                                                                  max_reads, min_score, min_HF, device=device,
                                                                  panel=kwargs.get('panel', None), fast_rng=fast_rng, _stats_out=stats)
        if not stats:                                   # nothing evaluated, or a window with one draw (statistics() raises there)
            stats = statistics(likelihoods, genomic_windows, device=device)
    some_statistics = {'matched_alleles': matched_alleles, 'runtime': time.time() - time0}
    info.update({'ancestral_makeup': ancestral_makeup, 'window_size': window_size, 'subsamples': subsamples, 'offset': offset,
                 'min_reads': min_reads, 'max_reads': max_reads, 'min_score': min_score, 'min_HF': min_HF,
                 'statistics': {**stats, **some_statistics}})
    if output_filename != None:
        save_results(likelihoods, info, compress, obs_filename, output_filename, kwargs.get('output_dir', 'results'))
    print_summary(obs_filename, info)
    print('Done calculating LLRs for all the genomic windows in %.3f sec.' % (time.time() - time0))
    return likelihoods, info


def parse_ancestral_makeup(tokens):
    """ The positional ANCESTRAL_MAKEUP of the CLI (ANEUPLOIDY_TEST.py:481-489): one or two population names -> set
    (non-admixed / recent admixture); NAME PROPORTION pairs -> dict (distant admixture). """
    names, rest = tokens[0::2], tokens[1::2]
    is_number = lambda t: t.replace('.', '', 1).isdigit()
    if all(t.isalpha() for t in tokens) and len(tokens) in (1, 2):
        return set(tokens)
    if all(t.isalpha() for t in names) and all(map(is_number, rest)):
        return dict(zip(names, map(float, rest)))
    raise Exception('error: Invalid argument supplied for ancestral makeup.')


def main(argv=None):
    """ The command line of ANEUPLOIDY_TEST.py (:440-493): same positionals, flags and defaults. """
    import argparse
    parser = argparse.ArgumentParser(description='Log-likelihood ratios of BPH / SPH / disomy / monosomy for the genomic windows of one '
                                                 'chromosome of a low-coverage sample, computed on a B200 GPU.')
    for name, text in (('obs_filename', 'observations file made by MAKE_OBS_TAB'), ('leg_filename', 'legend file of the reference panel'),
                       ('hap_filename', 'haplotype file of the reference panel')):
        parser.add_argument(name, metavar=name.upper(), type=str, help=text)
    parser.add_argument('ancestral_makeup', metavar='ANCESTRAL_MAKEUP', nargs='+',
                        help='EUR (non-admixed) | EUR EAS (recent admixture) | EUR 0.8 EAS 0.1 SAS 0.1 (distant admixture)')
    for short, name, kind, default, text in (
            ('-w', '--window-size', int, '100000', 'genomic window in bp; 0 = adaptive windows holding min-reads reads'),
            ('-s', '--subsamples', int, '32', 'bootstrap draws per genomic window'),
            ('-o', '--offset', int, 0, 'shift of all genomic windows in bp'),
            ('-m', '--min-reads', int, 6, 'windows with fewer scored reads are skipped'),
            ('-M', '--max-reads', int, 4, 'reads per bootstrap draw'),
            ('-F', '--min-HF', float, 0.05, 'haplotype frequencies in [FLOAT, 1-FLOAT] add to the score of a read'),
            ('-S', '--min-score', int, 2, 'minimal score of a read')):
        parser.add_argument(short, name, type=kind, metavar='FLOAT' if kind is float else 'INT', default=default, help=text)
    parser.add_argument('-O', '--output-filename', type=str, metavar='output_filename', default='',
                        help='default: the input name with .obs.p replaced by .LLR.p')
    parser.add_argument('-C', '--compress', metavar='gz/bz2/unc', type=str, default='unc', choices=['gz', 'bz2', 'unc'])
    args = vars(parser.parse_args(argv))
    args['ancestral_makeup'] = parse_ancestral_makeup(args['ancestral_makeup'])
    return aneuploidy_test(**args)
