"""
Host side of the likelihood engine: packs reference-panel rows into 64-bit
haplotype bit-planes, keeps them resident on one GPU and exposes the batched
operations of libldpgta.so (include/ldpgta.h) on numpy arrays.

One `Engine` = one (sample, chromosome) = the state the reference builds in
homogeneous.__init__ / recent_admixture.__init__ / distant_admixture.__init__
(HOMOGENOUES_MODELS.py:52-73 and twins).
"""
import ctypes

import numpy as np

from . import _native as N

LLR_PAIRS = (('BPH', 'SPH'), ('disomy', 'monosomy'), ('BPH', 'disomy'), ('disomy', 'SPH'), ('SPH', 'monosomy'))


def pack_ints(ints, n_bits):
    """ Python ints (bit i = haplotype i, as produced by MAKE_REF_PANEL.py:238,315)
    -> uint64[len, ceil(n_bits/64)], little-endian word order.  Values must fit
    in n_bits (TEST_MODULE.py:29-30 can emit 1 << N, which no panel contains). """
    W = (n_bits + 63) // 64
    nbytes = 8 * W
    if len(ints) and (min(ints) < 0 or max(ints) >> n_bits):
        bad = next(i for i, x in enumerate(ints) if x < 0 or x >> n_bits)
        raise ValueError('haplotype row %d does not fit in %d bits' % (bad, n_bits))
    return np.frombuffer(b''.join([x.to_bytes(nbytes, 'little') for x in ints]), dtype='<u8').reshape(len(ints), W)


def unpack_row(row):
    """ uint64[W] -> Python int. """
    return int.from_bytes(np.ascontiguousarray(row, dtype='<u8').tobytes(), 'little')


class ResidentPanel:
    """ A reference panel kept on one GPU for any number of samples: the LEGEND/HAP tables of one
    chromosome (MAKE_REF_PANEL.py:352-368) packed and uploaded once.  Model objects built with
    `panel=` gather their allele rows from it on the device instead of re-packing panel rows
    (build_hap_dict, HOMOGENOUES_MODELS.py:76-100). """

    def __init__(self, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, device=0, chunk=65536):
        self.lib = N.load()
        self.groups = list(hap_tab_per_group)
        self.n_hap = [int(number_of_haplotypes_per_group[g]) for g in self.groups]
        self.device = device
        self.n_snps = len(leg_tab)
        self.last_group = list(number_of_haplotypes_per_group)[-1]      # weights the read scores (ANEUPLOIDY_TEST.py:181-184)
        self._leg_tab, self._legend = leg_tab, None
        for g in self.groups:
            if len(hap_tab_per_group[g]) != self.n_snps:
                raise Exception('Error: the number of SNPs in the LEGEND file differ from the number of SNPs in the HAP file.')
        # position -> (ref, alt, legend row); a later duplicate position wins, as in the reference's dict
        self.by_pos = {leg[1]: (leg[2], leg[3], i) for i, leg in enumerate(leg_tab)}
        nh = np.asarray(self.n_hap, dtype=np.uint32)
        self._p = ctypes.c_void_p()
        N.check(self.lib.ldpgta_panel_create(device, len(self.groups), N.ptr(nh), self.n_snps, ctypes.byref(self._p)))
        for gi, g in enumerate(self.groups):
            tab = hap_tab_per_group[g]
            for lo in range(0, self.n_snps, chunk):
                rows = pack_ints(tab[lo:lo + chunk], self.n_hap[gi])
                N.check(self.lib.ldpgta_panel_upload(self._p, gi, lo, N.ptr(np.ascontiguousarray(rows)), rows.shape[0]))

    @property
    def legend(self):
        """ The legend as arrays (packed.PackedLegend), built on first use. """
        if self._legend is None:
            from .packed import PackedLegend
            self._legend = PackedLegend(self._leg_tab)
        return self._legend

    def close(self):
        if self._p is not None and self._p.value:
            self.lib.ldpgta_panel_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, haplotypes_per_group, device=0):
        self.lib = N.load()
        self.n_hap = [int(x) for x in haplotypes_per_group]
        self.G = len(self.n_hap)
        self.device = device
        nh = np.asarray(self.n_hap, dtype=np.uint32)
        self._ctx = ctypes.c_void_p()
        N.check(self.lib.ldpgta_create(device, self.G, N.ptr(nh), ctypes.byref(self._ctx)))
        self.words = [N.check(self.lib.ldpgta_row_words(self._ctx, g)) for g in range(self.G)]
        self.padded_words = [N.check(self.lib.ldpgta_padded_row_words(self._ctx, g)) for g in range(self.G)]
        self.n_alleles = 0
        self._keepalive = []
        self._plan = None

    # ---- lifetime -------------------------------------------------------
    def close(self):
        if self._ctx is not None and self._ctx.value:
            self.lib.ldpgta_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- panel residency --------------------------------------------------
    def upload_allele_rows(self, group, rows, complement=None):
        """ rows: uint64[n_alleles, words[group]] or a sequence of Python ints;
        complement: per-row flag, 1 = REF allele (row is complemented on device). """
        if not isinstance(rows, np.ndarray):
            rows = pack_ints(list(rows), self.n_hap[group])
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        if rows.ndim != 2 or rows.shape[1] != self.words[group]:
            raise ValueError('rows of group %d must have shape [n, %d]' % (group, self.words[group]))
        comp = None if complement is None else np.ascontiguousarray(complement, dtype=np.uint8)
        if comp is not None and comp.shape != (rows.shape[0],):
            raise ValueError('complement must have one flag per row')
        N.check(self.lib.ldpgta_upload_allele_rows(self._ctx, group, N.ptr(rows), rows.shape[0], N.ptr(comp)))
        self.n_alleles = rows.shape[0]

    def gather_allele_rows(self, panel, legend_idx, complement=None):
        """ Allele rows from a ResidentPanel: row a = panel row legend_idx[a] (complemented for REF alleles). """
        idx = np.ascontiguousarray(legend_idx, dtype=np.int64)
        comp = None if complement is None else np.ascontiguousarray(complement, dtype=np.uint8)
        N.check(self.lib.ldpgta_gather_allele_rows(self._ctx, panel._p, N.ptr(idx), N.ptr(comp), len(idx)))
        self.n_alleles = len(idx)

    def build_read_masks(self, allele_idx, read_offsets):
        idx = np.ascontiguousarray(allele_idx, dtype=np.int32)
        off = np.ascontiguousarray(read_offsets, dtype=np.int64)
        N.check(self.lib.ldpgta_build_read_masks(self._ctx, N.ptr(idx), N.ptr(off), len(off) - 1))

    def read_scores(self, allele_idx, read_offsets, alpha, score_weight, min_HF):
        """ build_score_dict (ANEUPLOIDY_TEST.py:151-189) for reads given as lists of allele rows. """
        idx = np.ascontiguousarray(allele_idx, dtype=np.int32)
        off = np.ascontiguousarray(read_offsets, dtype=np.int64)
        al = np.ascontiguousarray(alpha, dtype=np.float64)
        out = np.zeros(len(off) - 1, dtype=np.int32)
        N.check(self.lib.ldpgta_read_scores(self._ctx, N.ptr(idx), N.ptr(off), len(off) - 1, N.ptr(al),
                                            float(score_weight), float(min_HF), N.ptr(out)))
        return out

    def upload_read_masks(self, group, rows):
        if not isinstance(rows, np.ndarray):
            rows = pack_ints(list(rows), self.n_hap[group])
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        if rows.ndim != 2 or rows.shape[1] != self.words[group]:
            raise ValueError('rows of group %d must have shape [n, %d]' % (group, self.words[group]))
        N.check(self.lib.ldpgta_upload_read_masks(self._ctx, group, N.ptr(rows), rows.shape[0]))

    def attach_read_masks_dev(self, group, dev_ptr, n_rows, keepalive=None):
        """ Device-resident rows [n_rows][padded_words[group]] (e.g. a torch tensor's data_ptr). """
        N.check(self.lib.ldpgta_attach_read_masks_dev(self._ctx, group, ctypes.c_void_p(dev_ptr), n_rows))
        if keepalive is not None:
            self._keepalive.append(keepalive)

    @property
    def n_reads(self):
        return int(self.lib.ldpgta_num_reads(self._ctx))

    # ---- the hot path -------------------------------------------------------
    def joint_counts(self, read_idx):
        idx = np.ascontiguousarray(read_idx, dtype=np.int32)
        n = len(idx)
        out = np.zeros((self.G, 1 << n), dtype=np.uint32)
        N.check(self.lib.ldpgta_joint_counts(self._ctx, N.ptr(idx), n, N.ptr(out)))
        return out

    def get_likelihoods(self, kind, allele_lists, proportions=None, general=False, want_counts=False):
        """ One draw given as lists of allele-row indices (one list per haplotype). """
        n = len(allele_lists)
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(a) for a in allele_lists])
        idx = np.ascontiguousarray([i for a in allele_lists for i in a], dtype=np.int32)
        out = np.zeros(4, dtype=np.float64)
        counts = np.zeros((self.G, 1 << n), dtype=np.uint32) if want_counts and 2 <= n <= N.MAX_READS_PER_DRAW else None
        prop = None if proportions is None else np.ascontiguousarray(proportions, dtype=np.float64)
        N.check(self.lib.ldpgta_get_likelihoods(self._ctx, kind, N.ptr(idx), N.ptr(off), n, N.ptr(prop),
                                                1 if general else 0, N.ptr(out), N.ptr(counts)))
        return (out, counts) if want_counts else out

    def likelihoods_batch(self, kind, read_idx, draw_offsets, proportions=None, out=None):
        """ get_likelihoods for a batch of draws (the bootstrap loop ANEUPLOIDY_TEST.py:277-286).
        Returns float64[n_draws, 4] = monosomy, disomy, SPH, BPH. """
        idx = np.ascontiguousarray(read_idx, dtype=np.int32)
        off = np.ascontiguousarray(draw_offsets, dtype=np.int64)
        nd = len(off) - 1
        if out is None:
            out = np.empty((nd, 4), dtype=np.float64)
        prop = None if proportions is None else np.ascontiguousarray(proportions, dtype=np.float64)
        N.check(self.lib.ldpgta_likelihoods_batch(self._ctx, kind, N.ptr(idx), N.ptr(off), nd, N.ptr(prop), N.ptr(out)))
        return out

    def likelihoods_uniform(self, kind, read_idx, proportions=None, out=None):
        """ likelihoods_batch for draws of one size: read_idx[n_draws][n] (host; page-locked PinnedArray buffers
        stream through the copy/compute pipeline).  Returns float64[n_draws, 4]. """
        idx = np.ascontiguousarray(read_idx, dtype=np.int32)
        if idx.ndim != 2:
            raise ValueError('read_idx must be [n_draws][n]')
        nd, n = idx.shape
        if out is None:
            out = np.empty((nd, 4), dtype=np.float64)
        prop = None if proportions is None else np.ascontiguousarray(proportions, dtype=np.float64)
        N.check(self.lib.ldpgta_likelihoods_uniform(self._ctx, kind, n, N.ptr(idx), nd, N.ptr(prop), N.ptr(out)))
        return out

    def bootstrap_uniform(self, kind, n, window_offsets, window_reads, draw_offsets, seed, proportions=None, out=None,
                          want_draws=False):
        """ Draws made on the device (pick_reads for all windows at once) and evaluated: window w = read rows
        window_reads[window_offsets[w]:window_offsets[w+1]], draw_offsets[w+1]-draw_offsets[w] draws of n reads.
        Returns float64[n_draws, 4] (and the draws int32[n_draws, n] with want_draws). """
        woff = np.ascontiguousarray(window_offsets, dtype=np.int64)
        wr = np.ascontiguousarray(window_reads, dtype=np.int32)
        doff = np.ascontiguousarray(draw_offsets, dtype=np.int64)
        if len(woff) != len(doff):
            raise ValueError('window_offsets and draw_offsets must have the same length')
        nd = int(doff[-1])
        if out is None:
            out = np.empty((nd, 4), dtype=np.float64)
        idx = np.empty((nd, n), dtype=np.int32) if want_draws else None
        prop = None if proportions is None else np.ascontiguousarray(proportions, dtype=np.float64)
        N.check(self.lib.ldpgta_bootstrap_uniform(self._ctx, kind, n, N.ptr(woff), N.ptr(wr), len(woff) - 1, N.ptr(doff),
                                                  ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), N.ptr(prop), N.ptr(out),
                                                  N.ptr(idx) if want_draws else None))
        return (out, idx) if want_draws else out

    def likelihoods_uniform_dev(self, kind, n, idx_dev_ptr, n_draws, out_dev_ptr, proportions=None):
        """ Device-resident variant (asynchronous on the engine's stream). """
        prop = None if proportions is None else np.ascontiguousarray(proportions, dtype=np.float64)
        N.check(self.lib.ldpgta_likelihoods_uniform_dev(self._ctx, kind, n, ctypes.c_void_p(idx_dev_ptr), n_draws,
                                                        N.ptr(prop), ctypes.c_void_p(out_dev_ptr)))

    def synchronize(self):
        N.check(self.lib.ldpgta_synchronize(self._ctx))

    @property
    def stream(self):
        return int(self.lib.ldpgta_stream(self._ctx))

    @property
    def launch_count(self):
        return int(self.lib.ldpgta_launch_count(self._ctx))

    @property
    def graph_replays(self):
        """ host-pipeline calls served by one CUDA graph launch """
        return int(self.lib.ldpgta_graph_replays(self._ctx))

    # ---- bootstrap() + statistics() resident on the device ------------------------------------------
    def set_windows(self, window_offsets, window_reads, draw_offsets, reads_per_draw, segment_offsets=None,
                    segment_cols=None, segment_first=None):
        """ Makes the windows (build_gw_dict as CSR arrays), the draw plan (draws per window, reads per draw) and the
        (embryo, chromosome) segments resident; see ldpgta_set_windows.  window_reads=None: window w owns the read
        rows window_offsets[w]:window_offsets[w+1]; window_offsets=None: replay-only plan. """
        i64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int64)
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        woff, wr, doff, rpd = i64(window_offsets), i32(window_reads), i64(draw_offsets), i32(reads_per_draw)
        nw = len(doff) - 1
        if np.ndim(reads_per_draw) == 0:
            rpd = np.full(nw, int(reads_per_draw), dtype=np.int32)
        if len(rpd) != nw or (woff is not None and len(woff) != nw + 1):
            raise ValueError('window_offsets, draw_offsets and reads_per_draw disagree on the number of windows')
        soff, scols, sfirst = i64(segment_offsets), i32(segment_cols), i32(segment_first)
        ns = 1 if soff is None else len(soff) - 1
        for a in (scols, sfirst):
            if a is not None and len(a) != ns:
                raise ValueError('one value per segment is needed')
        N.check(self.lib.ldpgta_set_windows(self._ctx, N.ptr(woff), N.ptr(wr), nw, N.ptr(doff), N.ptr(rpd), N.ptr(soff), ns,
                                            N.ptr(scols), N.ptr(sfirst)))
        self._plan = self.plan_info()
        self._plan['reads_per_draw'] = rpd
        return self._plan

    def plan_info(self):
        out = np.zeros(10, dtype=np.int64)
        N.check(self.lib.ldpgta_plan_info(self._ctx, N.ptr(out)))
        keys = ('n_windows', 'n_draws', 'n_segments', 'ncp', 'n_classes', 'nnz', 'lik_dev', 'mean_var_dev', 'partials_dev', 'stats_dev')
        return dict(zip(keys, out.tolist()))

    def _stats_outputs(self, want, lik_out=None, out=None):
        p = self._plan if self._plan is not None else self.plan_info()      # raises without a resident plan
        nd, nw, ns, ncp = p['n_draws'], p['n_windows'], p['n_segments'], p['ncp']
        if out is not None:                      # caller-owned (e.g. page-locked) arrays, reused from call to call
            shapes = {'lik': (nd, 4), 'draws': (p['nnz'],), 'llr': (N.NUM_LLR_PAIRS, nd), 'mean_var': (N.NUM_LLR_PAIRS, nw, 2),
                      'partials': (ns, N.NUM_LLR_PAIRS, ncp + 3), 'stats': (ns, N.NUM_LLR_PAIRS, 3)}
            for k, a in out.items():
                if tuple(a.shape) != shapes[k] or a.dtype != (np.int32 if k == 'draws' else np.float64) or not a.flags['C_CONTIGUOUS']:
                    raise ValueError('output %r must be a C-contiguous array of shape %r' % (k, shapes[k]))
            return out
        out = {}
        if 'lik' in want:
            out['lik'] = lik_out if lik_out is not None else np.empty((nd, 4), dtype=np.float64)
        if 'draws' in want:
            out['draws'] = np.empty(p['nnz'], dtype=np.int32)
        if 'llr' in want:
            out['llr'] = np.empty((N.NUM_LLR_PAIRS, nd), dtype=np.float64)
        if 'mean_var' in want:
            out['mean_var'] = np.empty((N.NUM_LLR_PAIRS, nw, 2), dtype=np.float64)
        if 'partials' in want:
            out['partials'] = np.empty((ns, N.NUM_LLR_PAIRS, ncp + 3), dtype=np.float64)
        if 'stats' in want:
            out['stats'] = np.empty((ns, N.NUM_LLR_PAIRS, 3), dtype=np.float64)
        return out

    def bootstrap_stats(self, kind, seed, proportions=None, want=('mean_var', 'stats'), lik_out=None, out=None):
        """ bootstrap() + statistics() (ANEUPLOIDY_TEST.py:250-325) for the resident plan with the draws made on the
        device: only the seed goes up, only `want` comes back -- any of 'lik' [n_draws][4], 'draws' (concatenated read
        rows), 'llr' [5][n_draws], 'mean_var' [5][n_windows][2], 'partials' [n_segments][5][ncp+3],
        'stats' [n_segments][5][3] = mean_of_mean, std_of_mean, fraction_of_negative_LLRs.  `out` (a dict of arrays with
        those keys) replaces `want` and is filled in place. """
        out = self._stats_outputs(want, lik_out, out)
        prop = None if proportions is None else np.ascontiguousarray(proportions, dtype=np.float64)
        g = lambda k: N.ptr(out[k]) if k in out else None
        N.check(self.lib.ldpgta_bootstrap_stats(self._ctx, kind, ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), N.ptr(prop), g('lik'),
                                                g('draws'), g('llr'), g('mean_var'), g('partials'), g('stats')))
        return out

    def replay_stats(self, kind, read_idx, proportions=None, want=('lik', 'mean_var', 'stats'), lik_out=None, out=None):
        """ The same for draws given by the caller (read rows of every draw of the plan, window order). """
        idx = np.ascontiguousarray(read_idx, dtype=np.int32).reshape(-1)
        if len(idx) != self._plan['nnz']:
            raise ValueError('the plan holds %d read indices, got %d' % (self._plan['nnz'], len(idx)))
        out = self._stats_outputs([w for w in want if w != 'draws'], lik_out, out)
        prop = None if proportions is None else np.ascontiguousarray(proportions, dtype=np.float64)
        g = lambda k: N.ptr(out[k]) if k in out else None
        N.check(self.lib.ldpgta_replay_stats(self._ctx, kind, N.ptr(idx), N.ptr(prop), g('lik'), g('llr'), g('mean_var'),
                                             g('partials'), g('stats')))
        return out

    def bootstrap_stats_dev(self, kind, seed, proportions=None, partials_dev_ptr=None):
        """ Asynchronous, everything stays in HBM (see ldpgta_bootstrap_stats_dev). """
        prop = None if proportions is None else np.ascontiguousarray(proportions, dtype=np.float64)
        N.check(self.lib.ldpgta_bootstrap_stats_dev(self._ctx, kind, ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), N.ptr(prop),
                                                    ctypes.c_void_p(partials_dev_ptr) if partials_dev_ptr else None))

    def finish_segments_dev(self, partials_dev_ptr=None, stats_dev_ptr=None):
        N.check(self.lib.ldpgta_finish_segments_dev(self._ctx, ctypes.c_void_p(partials_dev_ptr) if partials_dev_ptr else None,
                                                    ctypes.c_void_p(stats_dev_ptr) if stats_dev_ptr else None))

    # ---- the cross-GPU exchange (peer memory over NVLink, comm_api.cu) -------------------------------------------
    def comm_create(self, rank, world, max_doubles):
        """ Allocates this rank's exchange buffer; returns its handle (bytes) for the peers. """
        h = np.zeros(N.COMM_HANDLE_BYTES, dtype=np.uint8)
        N.check(self.lib.ldpgta_comm_create(self._ctx, rank, world, int(max_doubles), N.ptr(h)))
        return h.tobytes()

    def comm_connect(self, handles):
        """ handles: the bytes objects of comm_create of every rank, in rank order. """
        blob = np.frombuffer(b''.join(handles), dtype=np.uint8).copy()
        N.check(self.lib.ldpgta_comm_connect(self._ctx, N.ptr(blob)))

    def allreduce_chrom_dev(self, partials_dev_ptr, n_doubles):
        """ In-place sum over ranks of a device buffer of float64 (asynchronous on the engine's stream). """
        N.check(self.lib.ldpgta_allreduce_chrom_dev(self._ctx, ctypes.c_void_p(partials_dev_ptr), int(n_doubles)))

    def comm_destroy(self):
        N.check(self.lib.ldpgta_comm_destroy(self._ctx))

    def kernel_timing(self, enable=True, stride=1):
        """ CUDA-event brackets around the draw kernels of bootstrap_stats / replay_stats (measurement aid);
        enable=2: around the reductions of bootstrap_stats_dev instead, 3: its window-statistics kernel, 4: its sampler.
        stride=k records only every k-th bracket (an event pair costs the stream about 5 us). """
        mode = int(enable) if enable in (2, 3, 4) else (1 if enable else 0)
        N.check(self.lib.ldpgta_kernel_timing(self._ctx, mode | (max(1, int(stride)) << 8 if mode else 0)))

    def kernel_times(self, max_entries=65536):
        """ milliseconds of every bracket since the last call (waits for the stream) """
        out = np.zeros(max_entries, dtype=np.float64)
        n = N.check(self.lib.ldpgta_kernel_times(self._ctx, N.ptr(out), max_entries))
        return out[:min(n, max_entries)]

    # ---- reductions -----------------------------------------------------------
    def window_stats(self, lik, window_offsets, n_cols, want_llr=True):
        """ LLRs, per-window mean / sample variance and the chromosome partial sums
        (statistics(), ANEUPLOIDY_TEST.py:289-325) for consecutive windows of draws.  The per-draw LLRs also stay
        on the device for bin_stats(None, ...); want_llr=False skips their download. """
        lik = np.ascontiguousarray(lik, dtype=np.float64)
        off = np.ascontiguousarray(window_offsets, dtype=np.int64)
        nw, nd = len(off) - 1, lik.shape[0]
        llr = np.empty((N.NUM_LLR_PAIRS, nd), dtype=np.float64) if want_llr else None
        mean_var = np.empty((N.NUM_LLR_PAIRS, nw, 2), dtype=np.float64)
        partials = np.empty((N.NUM_LLR_PAIRS, n_cols + 3), dtype=np.float64)
        N.check(self.lib.ldpgta_window_stats(self._ctx, N.ptr(lik), N.ptr(off), nw, n_cols, N.ptr(llr) if want_llr else None, N.ptr(mean_var),
                                             N.ptr(partials)))
        return llr, mean_var, partials

    def bin_stats(self, x, window_offsets, bin_ranges):
        """ Mean LLR and its standard error per bin of windows (binning(), PLOT_PANEL.py:121-144).
        x: [n_series][n_draws] values, or None = the five LLR series the last window_stats() call left on the
        device.  bin_ranges: [n_bins][2] window index ranges.  Returns [n_series][n_bins][2] = (mean, std). """
        off = np.ascontiguousarray(window_offsets, dtype=np.int64)
        rng = np.ascontiguousarray(bin_ranges, dtype=np.int64).reshape(-1, 2)
        if x is None:
            ns, xp = N.NUM_LLR_PAIRS, None
        else:
            x = np.ascontiguousarray(x, dtype=np.float64)
            x = x.reshape(1, -1) if x.ndim == 1 else x
            if x.shape[1] != off[-1]:
                raise ValueError('x has %d values per series, the windows hold %d draws' % (x.shape[1], off[-1]))
            ns, xp = x.shape[0], N.ptr(x)
        out = np.empty((ns, len(rng), 2), dtype=np.float64)
        N.check(self.lib.ldpgta_bin_stats(self._ctx, xp, ns, N.ptr(off), len(off) - 1, N.ptr(rng), len(rng), N.ptr(out)))
        return out


def finish_chromosome(partials, n_cols, first_window_draws):
    """ mean_of_mean, std_of_mean, fraction_of_negative_LLRs per LLR pair from
    (summed) chromosome partials -- ANEUPLOIDY_TEST.py:65-76, 311-313. """
    partials = np.ascontiguousarray(partials, dtype=np.float64)
    out = np.empty((N.NUM_LLR_PAIRS, 3), dtype=np.float64)
    N.check(N.load().ldpgta_finish_chromosome(N.ptr(partials), n_cols, first_window_draws, N.ptr(out)))
    return out
