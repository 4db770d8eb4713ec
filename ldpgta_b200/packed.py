"""
packed -- the array-native route through one (sample, chromosome): SURVEY 8(f) rows 2 and 3.

The reference turns the observation table into dicts of Python objects before the first likelihood is computed
(build_hap_dict HOMOGENOUES_MODELS.py:76-100, build_reads_dict / build_overlaps_dict / build_score_dict /
build_gw_dict ANEUPLOIDY_TEST.py:118-224); on a GPU those dict stages cost a thousand times more than the
likelihoods.  Here the same quantities are index arrays:

    PackedLegend     legend positions / ref / alt as arrays (once per panel)
    PackedSample     obs_tab matched against a ResidentPanel: allele table (legend row + complement flag, gathered
                     on the device), reads as CSR lists of allele rows (read masks built on the device), read scores
                     from the device, fraction_of_matches
    windows()        build_gw_dict as arrays: window borders + CSR of read rows, fixed or adaptive, with the
                     reference's quirks (offset, adaptive growth in 10 kbp steps, last window dropped, empty
                     windows kept)
    plan()           all bootstrap draws of all windows by vectorised sampling without replacement on the host, or
    bootstrap_on_device()  the draws made on the device (ldpgta_bootstrap_uniform, Philox + Floyd): only window lists go up
    run()            draws -> likelihoods on the device -> LLR statistics on the device -> PackedResult, which can
                     also be unfolded into the reference's (likelihoods, genomic_windows, fraction_of_matches)

Draws come from a numpy Generator, not from random.sample on hash-ordered tuples, so a packed run is statistically
equivalent to the reference's, not stream-identical (the drop-in ANEUPLOIDY_TEST.bootstrap keeps the stream).
Everything else is pinned to the golden vectors in tests/test_packed.py (read scores, windows, fraction of matches,
and likelihoods for replayed draws).
"""
from math import comb

import numpy as np

from . import _native as N
from .engine import Engine, LLR_PAIRS, finish_chromosome


class PackedLegend:
    """ LEGEND rows (chr_id, pos, ref, alt) as arrays.  Lookup keeps the reference's dict semantics: of several rows
    with one position the last wins (HOMOGENOUES_MODELS.py:84). """

    def __init__(self, leg_tab):
        n = len(leg_tab)
        self.codes = {}
        code = self.codes.setdefault
        pos = np.fromiter((l[1] for l in leg_tab), dtype=np.int64, count=n)
        ref = np.fromiter((code(l[2], len(self.codes)) for l in leg_tab), dtype=np.int32, count=n)
        alt = np.fromiter((code(l[3], len(self.codes)) for l in leg_tab), dtype=np.int32, count=n)
        order = np.argsort(pos, kind='stable')
        self.pos, self.ref, self.alt, self.row = pos[order], ref[order], alt[order], order.astype(np.int64)
        self.n_snps = n

    def encode_bases(self, bases, count):
        """ base strings -> codes; strings the legend never uses get -1 (they can not match). """
        get = self.codes.get
        return np.array([get(b, -1) for b in bases], dtype=np.int32).reshape(count)

    def match(self, pos, base):
        """ (keep, legend row, complement flag) per observation: an observed base that equals ALT selects the panel
        row, one that equals REF its complement, anything else is a mismatch (HOMOGENOUES_MODELS.py:89-96). """
        j = np.searchsorted(self.pos, pos, side='right') - 1
        hit = (j >= 0) & (self.pos[np.maximum(j, 0)] == pos)
        j = np.where(hit, j, 0)
        is_alt = hit & (base == self.alt[j]) & (base >= 0)
        is_ref = hit & ~is_alt & (base == self.ref[j]) & (base >= 0)
        return is_alt | is_ref, self.row[j], is_ref


def sample_without_replacement(rng, R, k):
    """ k distinct integers of range(R[d]) for every draw d (Floyd's algorithm, vectorised over draws);
    uniform over k-subsets, order inside a draw arbitrary (the likelihoods are symmetric in the reads). """
    R = np.asarray(R, dtype=np.int64)
    out = np.empty((len(R), k), dtype=np.int64)
    for i in range(k):
        top = R - k + i                                   # Floyd: t uniform in [0, top], take top instead if already chosen
        t = (rng.random(len(R)) * (top + 1)).astype(np.int64)
        np.minimum(t, top, out=t)
        if i:
            dup = (out[:, :i] == t[:, None]).any(axis=1)
            t = np.where(dup, top, t)
        out[:, i] = t
    return out


class PinnedPool:
    """ Page-locked result buffers, reused across samples (allocating page-locked memory costs far more than a run). """

    def __init__(self):
        self.free = []

    def take(self, n_draws):
        need = max(1, int(n_draws)) * 4
        for i, buf in enumerate(self.free):
            if buf.array.size >= need:
                return self.free.pop(i)
        return N.PinnedArray((need + need // 4,), np.float64)

    def give(self, buf):
        if buf is not None and len(self.free) < 4:
            self.free.append(buf)


class PackedWindows:
    """ Genomic windows of one chromosome: borders[K][2] = (a, b-1) as in gw_dict's keys, CSR offsets / read rows. """

    def __init__(self, borders, offsets, reads):
        self.borders, self.offsets, self.reads = borders, offsets, reads

    def __len__(self):
        return len(self.borders)

    @property
    def sizes(self):
        return np.diff(self.offsets)


class PackedResult:
    """ Likelihoods of all draws of a chromosome, window-major: lik[draw_offsets[w]:draw_offsets[w+1]] belong to
    evaluated window w (borders[w]); statistics as computed on the device. """

    def __init__(self, sample, windows, evaluated, draw_offsets, draws, lik, mean_var, chrom, lease=None):
        self.sample, self.windows, self.evaluated = sample, windows, evaluated
        self.draw_offsets, self.draws, self.likelihoods = draw_offsets, draws, lik
        self.window_mean_var, self.chromosome = mean_var, chrom
        self._lease = lease                 # (pool, page-locked buffer) behind `likelihoods`

    def release(self):
        """ hands the page-locked buffer behind `likelihoods` back for the next sample (the array becomes invalid) """
        if self._lease is not None:
            pool, buf = self._lease
            self._lease, self.likelihoods = None, None
            pool.give(buf)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    @property
    def borders(self):
        return self.windows.borders[self.evaluated]

    def llr_per_chromosome(self):
        """ {pair: {'mean_of_mean', 'std_of_mean', 'fraction_of_negative_LLRs'}} as in statistics() (:311-313). """
        return {pair: {'mean_of_mean': float(self.chromosome[p, 0]), 'std_of_mean': float(self.chromosome[p, 1]),
                       'fraction_of_negative_LLRs': float(self.chromosome[p, 2])} for p, pair in enumerate(LLR_PAIRS)}

    def as_reference(self, tuple_type=None):
        """ (likelihoods, genomic_windows, fraction_of_matches) in the reference's structures (ANEUPLOIDY_TEST.py:287). """
        if tuple_type is None:
            from .models import likelihoods_tuple as tuple_type
        names = self.sample.read_names
        gw = {(int(a), int(b)): tuple(names[r] for r in self.windows.reads[lo:hi])
              for (a, b), lo, hi in zip(self.windows.borders.tolist(), self.windows.offsets[:-1].tolist(), self.windows.offsets[1:].tolist())}
        rows = self.likelihoods.tolist()
        lik = {(int(a), int(b)): tuple(tuple_type(*rows[i]) for i in range(lo, hi))
               for (a, b), lo, hi in zip(self.borders.tolist(), self.draw_offsets[:-1].tolist(), self.draw_offsets[1:].tolist())}
        return lik, gw, dict(self.sample.fraction_of_matches)


class PackedObservations:
    """ obs_tab rows that match the legend, as arrays (see pack_observations).  save() / load() keep them as an .npz
    next to the sample so that later runs against the same panel skip the walk over the pickled tuples. """
    _arrays = ('allele_rows', 'allele_complement', 'read_offsets', 'read_alleles', 'obs_pos', 'obs_read')

    def save(self, filename):
        np.savez(filename, read_names=np.array(self.read_names, dtype=str), fraction_of_matches=np.float64(self.fraction_of_matches),
                 n_snps=np.int64(self.n_snps), **{k: getattr(self, k) for k in self._arrays})

    @classmethod
    def load(cls, filename):
        out = cls()
        with np.load(filename) as z:
            for k in cls._arrays:
                setattr(out, k, z[k])
            out.read_names = z['read_names'].tolist()
            out.fraction_of_matches = float(z['fraction_of_matches'])
            out.n_snps = int(z['n_snps'])
        return out


def pack_observations(obs_tab, legend):
    """ obs_tab [(pos, read_id, base)] against a PackedLegend:
      fraction_of_matches            1 - mismatches / len(obs_tab)                       (HOMOGENOUES_MODELS.py:98)
      allele_rows, allele_complement the distinct observed alleles = keys of hap_dict     (:89-94)
      read_names, read_offsets, read_alleles   reads as CSR lists of allele-table rows    (build_reads_dict, ANEUPLOIDY_TEST.py:118-127)
      obs_pos, obs_read              position and read row of every matched observation  (build_overlaps_dict, :129-139) """
    out = PackedObservations()
    out.n_snps = legend.n_snps
    n = len(obs_tab)
    cols = list(zip(*obs_tab)) if n else ((), (), ())        # one C-level pass over the tuples
    pos = np.array(cols[0], dtype=np.int64)
    names = dict.fromkeys(cols[1])                            # read ids in order of first appearance
    for i, k in enumerate(names):
        names[k] = i
    read = np.array(list(map(names.__getitem__, cols[1])), dtype=np.int64)
    base = legend.encode_bases(cols[2], n)
    keep, row, is_ref = legend.match(pos, base)
    out.fraction_of_matches = 1 - int((~keep).sum()) / n
    # allele table: one row per distinct (legend row, ref/alt) among the matched observations
    key = row[keep] * 2 + is_ref[keep]
    alleles, allele_of_obs = np.unique(key, return_inverse=True)
    out.allele_rows, out.allele_complement = alleles >> 1, (alleles & 1).astype(np.uint8)
    # reads: CSR over allele rows, observation order kept inside a read
    kread = read[keep]
    order = np.argsort(kread, kind='stable')
    codes, counts = np.unique(kread, return_counts=True)
    all_names = list(names)
    out.read_names = [all_names[c] for c in codes.tolist()]
    out.read_offsets = np.zeros(len(codes) + 1, dtype=np.int64)
    np.cumsum(counts, out=out.read_offsets[1:])
    out.read_alleles = allele_of_obs[order].astype(np.int32)
    out.obs_pos = pos[keep]
    out.obs_read = np.searchsorted(codes, kread)
    if len(out.obs_pos) > 1 and (np.diff(out.obs_pos) < 0).any():
        raise ValueError('the observation table must be sorted by position')
    return out


class PackedSample:
    """ One sample on one chromosome against a resident panel; obs_tab is the reference's list of (pos, read_id, base)
    rows or a PackedObservations made from it earlier. """

    def __init__(self, obs_tab, panel, ancestral_makeup, min_HF=0.05, device=None):
        self.panel = panel
        groups = panel.groups
        if type(ancestral_makeup) == dict:
            if list(ancestral_makeup) != groups:
                raise ValueError('distant admixture: build the ResidentPanel with the groups in the order of the ancestral makeup')
            if sum(ancestral_makeup.values()) < 0.999:
                raise Exception('Error: ancestry proportions must sum to one.')
            self.kind, self.proportions = N.DISTANT_ADMIXTURE, [float(ancestral_makeup[g]) for g in groups]
            makeup = dict(ancestral_makeup)
        else:
            if set(ancestral_makeup) != set(groups) or len(groups) not in (1, 2):
                raise ValueError('the ancestral makeup does not match the superpopulations of the resident panel')
            self.kind, self.proportions = (N.HOMOGENEOUS if len(groups) == 1 else N.RECENT_ADMIXTURE), None
            makeup = {g: 1 / len(groups) for g in groups}
        if isinstance(obs_tab, PackedObservations):
            packed = obs_tab
            if packed.n_snps != panel.n_snps or (len(packed.allele_rows) and int(packed.allele_rows.max()) >= panel.n_snps):
                raise ValueError('these packed observations were matched against another panel')
        else:
            packed = pack_observations(obs_tab, panel.legend)
        frac = packed.fraction_of_matches
        self.fraction_of_matches = {g: frac for g in groups}
        self.allele_rows, self.allele_complement = packed.allele_rows, packed.allele_complement
        self.read_names, self.read_offsets, self.read_alleles = packed.read_names, packed.read_offsets, packed.read_alleles
        self.obs_pos, self.obs_read = packed.obs_pos, packed.obs_read
        if not hasattr(panel, '_result_pool'):
            panel._result_pool = PinnedPool()
        self.pool = panel._result_pool
        self.engine = Engine(panel.n_hap, device=panel.device if device is None else device)
        self.engine.gather_allele_rows(panel, self.allele_rows, self.allele_complement)
        self.engine.build_read_masks(self.read_alleles, self.read_offsets)
        alpha = [makeup[g] / nh for g, nh in zip(groups, panel.n_hap)]
        weight = makeup[panel.last_group] / panel.n_hap[groups.index(panel.last_group)]      # last-alpha quirk (:181-184)
        self.scores = self.engine.read_scores(self.read_alleles, self.read_offsets, alpha, weight, min_HF)
        self._lease = None

    def close(self):
        self.engine.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- build_gw_dict (ANEUPLOIDY_TEST.py:191-224) -----------------------------------------------------------
    def windows(self, window_size, offset, min_reads, max_reads, min_score, scores=None):
        return build_windows(self.obs_pos, self.obs_read, self.scores if scores is None else scores,
                             window_size, offset, min_reads, min_score)

    # ---- pick_reads / effective_number_of_subsamples (:226-248) for all windows at once ---------------------------
    @staticmethod
    def plan_counts(windows, min_reads, max_reads, subsamples):
        """ (evaluated windows, draw_offsets over them, draws per window, reads per draw per window). """
        R = windows.sizes
        effN = np.zeros(len(R), dtype=np.int64)
        big = (R >= min_reads) & (R > max_reads)
        small = (R >= min_reads) & (R <= max_reads)
        # comb(R, max_reads) < subsamples only for R barely above max_reads: evaluate exactly below that bound
        effN[big] = subsamples
        bound = max_reads + 1
        while comb(bound, max_reads) < subsamples:
            bound += 1
        for w in np.flatnonzero(big & (R < bound)).tolist():
            effN[w] = comb(int(R[w]), max_reads)
        effN[small] = np.minimum(R[small], subsamples)
        evaluated = np.flatnonzero(effN > 0)
        draw_offsets = np.zeros(len(evaluated) + 1, dtype=np.int64)
        np.cumsum(effN[evaluated], out=draw_offsets[1:])
        return evaluated, draw_offsets, effN, np.minimum(R - 1, max_reads)

    @classmethod
    def plan(cls, windows, min_reads, max_reads, subsamples, rng):
        """ (evaluated windows, draw_offsets, [(draw rows, read_idx[n_draws_k][k]) per draw size k]), draws sampled on the host. """
        evaluated, draw_offsets, effN, k_of = cls.plan_counts(windows, min_reads, max_reads, subsamples)
        R = windows.sizes
        groups = []
        win_of_draw = np.repeat(evaluated, effN[evaluated])
        k_draw = k_of[win_of_draw]
        for k in np.unique(k_draw).tolist():
            rows = np.flatnonzero(k_draw == k)
            w = win_of_draw[rows]
            local = sample_without_replacement(rng, R[w], k)
            idx = windows.reads[windows.offsets[w][:, None] + local].astype(np.int32)
            groups.append((rows, idx))
        return evaluated, draw_offsets, groups

    def bootstrap_on_device(self, windows, min_reads, max_reads, subsamples, seed, want_draws=False):
        """ plan + evaluate with the draws made on the device (ldpgta_bootstrap_uniform): only the window lists go up.
        Returns (evaluated windows, draw_offsets, draw groups or None, likelihoods[n_draws][4]). """
        evaluated, draw_offsets, effN, k_of = self.plan_counts(windows, min_reads, max_reads, subsamples)
        n_draws = int(draw_offsets[-1])
        buf = self.pool.take(n_draws)
        lik = buf.array[:4 * n_draws].reshape(n_draws, 4)
        groups = [] if want_draws else None
        sizes = np.unique(k_of[evaluated]).tolist()
        for k in sizes:
            if self.kind == N.DISTANT_ADMIXTURE and k == 5:
                self.pool.give(buf)
                raise AttributeError("'distant_admixture' object has no attribute 'likelihoods5'")
            per_window = np.where(k_of == k, effN, 0)                    # windows of other draw sizes get no draws in this call
            doff = np.zeros(len(per_window) + 1, dtype=np.int64)
            np.cumsum(per_window, out=doff[1:])
            if len(sizes) == 1:                                          # the usual case: results land where they stay
                got = self.engine.bootstrap_uniform(self.kind, k, windows.offsets, windows.reads, doff, seed + k,
                                                    proportions=self.proportions, out=lik, want_draws=want_draws)
                if want_draws:
                    groups.append((np.arange(n_draws), got[1]))
                continue
            got = self.engine.bootstrap_uniform(self.kind, k, windows.offsets, windows.reads, doff, seed + k,
                                                proportions=self.proportions, want_draws=want_draws)
            sel = np.flatnonzero(k_of[evaluated] == k)                    # evaluated windows of this size, in order
            rows = np.concatenate([np.arange(draw_offsets[i], draw_offsets[i + 1]) for i in sel.tolist()]) if len(sel) else np.empty(0, np.int64)
            if want_draws:
                lik[rows] = got[0]
                groups.append((rows, got[1]))
            else:
                lik[rows] = got
        self._lease = (self.pool, buf)
        return evaluated, draw_offsets, groups, lik

    def evaluate(self, draw_groups, n_draws):
        """ likelihoods [n_draws][4] for groups of uniform draws (rows of the output, read_idx[.][k]). """
        buf = self.pool.take(n_draws)
        lik = buf.array[:4 * n_draws].reshape(n_draws, 4)
        for rows, idx in draw_groups:
            if self.kind == N.DISTANT_ADMIXTURE and idx.shape[1] == 5:
                self.pool.give(buf)
                raise AttributeError("'distant_admixture' object has no attribute 'likelihoods5'")
            if len(draw_groups) == 1 and len(rows) == n_draws:
                self.engine.likelihoods_uniform(self.kind, idx, proportions=self.proportions, out=lik)
            else:
                lik[rows] = self.engine.likelihoods_uniform(self.kind, idx, proportions=self.proportions)
        self._lease = (self.pool, buf)
        return lik

    def run(self, window_size=100000, subsamples=32, offset=0, min_reads=6, max_reads=4, min_score=2, rng=None, seed=None,
            want_draws=True, keep_pinned=False, replay=None, want_likelihoods=True):
        """ bootstrap() + statistics() (ANEUPLOIDY_TEST.py:250-325) on arrays, as ONE device-resident pipeline
        (ldpgta_set_windows + ldpgta_bootstrap_stats / ldpgta_replay_stats): the likelihoods go from the draw kernels to the
        LLR / window / chromosome reductions inside HBM.  The draws come from the device-side sampler when `seed` is
        given, from `replay` (read rows of every draw, window order: the reference's recorded random.sample stream), or
        from `rng` (a numpy Generator, sampled on the host).  want_likelihoods=False downloads only the statistics
        (80 bytes per window); keep_pinned=True leaves `likelihoods` in the page-locked buffer it was downloaded into (no
        copy), valid until the result's release() or deletion, after which the buffer serves the next sample. """
        win = self.windows(window_size, offset, min_reads, max_reads, min_score)
        evaluated, draw_offsets, effN, k_of = self.plan_counts(win, min_reads, max_reads, subsamples)
        if not len(evaluated):
            return PackedResult(self, win, evaluated, draw_offsets, [] if want_draws else None,
                                np.empty((0, 4)) if want_likelihoods else None, None, None, None)
        sizes = k_of[evaluated].astype(np.int32)
        if self.kind == N.DISTANT_ADMIXTURE and (sizes == 5).any():
            raise AttributeError("'distant_admixture' object has no attribute 'likelihoods5'")
        # read lists of the evaluated windows only
        counts = win.sizes[evaluated]
        woff = np.zeros(len(evaluated) + 1, dtype=np.int64)
        np.cumsum(counts, out=woff[1:])
        if len(evaluated) == len(win):
            wreads = win.reads
        else:
            take = np.repeat(win.offsets[:-1][evaluated] - woff[:-1], counts) + np.arange(woff[-1])
            wreads = win.reads[take]
        self.engine.set_windows(woff, wreads, draw_offsets, sizes)
        n_draws = int(draw_offsets[-1])
        want = ['mean_var', 'stats']
        buf = lik = None
        if want_likelihoods:
            buf = self.pool.take(n_draws)
            lik = buf.array[:4 * n_draws].reshape(n_draws, 4)
            want.append('lik')
        flat = None
        try:
            if seed is not None:
                if want_draws:
                    want.append('draws')
                out = self.engine.bootstrap_stats(self.kind, int(seed), proportions=self.proportions, want=want, lik_out=lik)
                flat = out.get('draws')
            else:
                if replay is not None:
                    flat = np.ascontiguousarray(replay, dtype=np.int32).reshape(-1)
                else:
                    rng = np.random.default_rng() if rng is None else rng
                    per_draw_k = np.repeat(sizes, np.diff(draw_offsets))
                    starts = np.zeros(n_draws + 1, dtype=np.int64)
                    np.cumsum(per_draw_k, out=starts[1:])
                    flat = np.empty(int(starts[-1]), dtype=np.int32)
                    for rows, idx in self.plan(win, min_reads, max_reads, subsamples, rng)[2]:
                        flat[starts[rows][:, None] + np.arange(idx.shape[1])] = idx
                out = self.engine.replay_stats(self.kind, flat, proportions=self.proportions, want=want, lik_out=lik)
        except Exception:
            self.pool.give(buf)
            raise
        groups = None
        if want_draws and flat is not None:
            per_draw_k = np.repeat(sizes, np.diff(draw_offsets))
            starts = np.zeros(n_draws + 1, dtype=np.int64)
            np.cumsum(per_draw_k, out=starts[1:])
            groups = []
            for k in np.unique(per_draw_k).tolist():
                rows = np.flatnonzero(per_draw_k == k)
                groups.append((rows, flat[starts[rows][:, None] + np.arange(k)]))
        lease = (self.pool, buf) if buf is not None else None
        if lik is not None and not keep_pinned:
            lik = np.array(lik)
            self.pool.give(buf)
            lease = None
        return PackedResult(self, win, evaluated, draw_offsets, groups, lik, out['mean_var'], out['stats'][0], lease)


def build_windows(obs_pos, obs_read, scores, window_size, offset, min_reads, min_score):
    """ build_gw_dict (ANEUPLOIDY_TEST.py:191-224) on sorted observation arrays.  A window is emitted when a later
    position closes it, so the chromosome's last window never appears; windows without qualifying reads are kept. """
    max_win_size, initial = 350000, 10000
    adaptive = not window_size
    ws = initial if adaptive else int(window_size)
    offset = int(offset)
    assert len(obs_pos), 'error: the observation table is empty.'
    first, last = int(obs_pos[0]) + offset, int(obs_pos[-1]) + ws
    ok = np.asarray(scores)[obs_read] >= min_score
    if not adaptive:
        K = max(0, (int(obs_pos[-1]) - first) // ws)                # windows closed by a later position
        k = (obs_pos - first) // ws
        sel = ok & (obs_pos >= first) & (k < K)
        stride = max(1, len(scores))
        pair = np.unique(k[sel] * stride + obs_read[sel])                  # distinct (window, read), window-major
        wk, reads = pair // stride, pair % stride
        a = first + ws * np.arange(K, dtype=np.int64)
        borders = np.stack([a, a + ws - 1], axis=1) if K else np.empty((0, 2), dtype=np.int64)
        offsets = np.zeros(K + 1, dtype=np.int64)
        np.cumsum(np.bincount(wk, minlength=K), out=offsets[1:])
        return PackedWindows(borders, offsets, reads.astype(np.int64))
    # adaptive windows: sequential in the window borders, vectorised inside a window
    borders, chunks, sizes = [], [], []
    a, b = first, first + ws
    n = len(obs_pos)
    while b < last:
        lo, hi = np.searchsorted(obs_pos, a, side='left'), np.searchsorted(obs_pos, b, side='left')
        if hi >= n:
            break                                                    # nothing beyond this window closes it
        members = np.unique(obs_read[lo:hi][ok[lo:hi]])
        if 0 < len(members) < min_reads and b - a <= max_win_size:
            b += 10000
            continue
        borders.append((a, b - 1))
        chunks.append(members)
        sizes.append(len(members))
        a, b = b, b + ws
    offsets = np.zeros(len(borders) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    reads = np.concatenate(chunks).astype(np.int64) if chunks else np.empty(0, dtype=np.int64)
    return PackedWindows(np.asarray(borders, dtype=np.int64).reshape(-1, 2), offsets, reads)
