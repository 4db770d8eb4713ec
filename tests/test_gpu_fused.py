"""
The device-resident bootstrap -> statistics pipeline (ldpgta_set_windows, ldpgta_bootstrap_stats, ldpgta_replay_stats,
bootstrap_api.cu) against (i) the unfused C-ABI calls it replaces (likelihoods bit-identical), (ii) the oracle's
statistics() = ANEUPLOIDY_TEST.py:289-325 on the same likelihoods (window mean/variance <= 1e-12, chromosome-level
mean / std <= 1e-9 as north_star states), with several (embryo, chromosome) segments in one context, mixed draw sizes,
windows given as read lists or as consecutive rows, and many draws per window.
"""
import numpy as np
import pytest

import ldpgta_oracle as O

pytestmark = pytest.mark.gpu

L = pytest.importorskip('ldpgta_b200')

STAT_TOL = 1e-9          # chromosome-level LLR mean / std (north_star)


def random_rows(rng, n_rows, n_hap, density=0.7):
    """ uint64[n_rows, ceil(n_hap / 64)] with bits above n_hap clear """
    words = (n_hap + 63) // 64
    bits = rng.random((n_rows, words * 64)) < density
    bits[:, n_hap:] = False
    return (bits.reshape(n_rows, words, 64).astype(np.uint64) << np.arange(64, dtype=np.uint64)).sum(axis=2, dtype=np.uint64)


def oracle_stats(lik, doff, seg_off):
    """ statistics() of the oracle per segment: [(per-window {pair: [(mean, var)]}, per-chromosome {pair: dict})]. """
    out = []
    for s in range(len(seg_off) - 1):
        wins = range(seg_off[s], seg_off[s + 1])
        likd = {(w, w): tuple(O.likelihoods_tuple(*lik[i]) for i in range(doff[w], doff[w + 1])) for w in wins}
        out.append(O.statistics(likd, {k: tuple(range(10)) for k in likd}))
    return out


def check_against_oracle(res, lik, doff, seg_off):
    want = oracle_stats(lik, doff, seg_off)
    mv, stats = res['mean_var'], res['stats']
    for s, st in enumerate(want):
        for p, pair in enumerate(L.LLR_PAIRS):
            for w in range(seg_off[s], seg_off[s + 1]):
                m, v = st['LLRs_per_genomic_window'][pair][(w, w)]
                assert abs(mv[p, w, 0] - m) <= 1e-13 * max(1, abs(m)), (s, pair, w)
                assert abs(mv[p, w, 1] - v) <= 1e-12 * max(1, abs(v)), (s, pair, w, mv[p, w, 1], v)
            c = st['LLRs_per_chromosome'][pair]
            assert abs(stats[s, p, 0] - c['mean_of_mean']) <= STAT_TOL * max(1e-3, abs(c['mean_of_mean']))
            assert abs(stats[s, p, 1] - c['std_of_mean']) <= STAT_TOL * max(1e-3, abs(c['std_of_mean']))
            assert stats[s, p, 2] == c['fraction_of_negative_LLRs']


def make_case(rng, n_windows, reads_lo, reads_hi, n_hap=331, shuffle=True):
    R = rng.integers(reads_lo, reads_hi, size=n_windows)
    n_reads = int(R.sum())
    rows = random_rows(rng, n_reads, n_hap)
    woff = np.concatenate([[0], np.cumsum(R)]).astype(np.int64)
    wreads = np.arange(n_reads, dtype=np.int32)
    if shuffle:                                   # read lists in arbitrary row order, shared rows between neighbours
        wreads = rng.permutation(n_reads).astype(np.int32)
    return rows, woff, wreads, R


def test_fused_equals_unfused_and_oracle_uniform():
    rng = np.random.default_rng(11)
    n_hap, n, subs = 331, 4, 32
    rows, woff, wreads, R = make_case(rng, 57, 6, 40)
    doff = np.arange(0, (len(R) + 1) * subs, subs, dtype=np.int64)
    seg_off = np.array([0, 20, 21, 57], dtype=np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        old_lik, old_idx = eng.bootstrap_uniform(L.HOMOGENEOUS, n, woff, wreads, doff, 99 + n, want_draws=True)
        eng.set_windows(woff, wreads, doff, n, segment_offsets=seg_off)
        res = eng.bootstrap_stats(L.HOMOGENEOUS, 99, want=('lik', 'draws', 'llr', 'mean_var', 'partials', 'stats'))
        assert np.array_equal(res['lik'], old_lik)
        assert np.array_equal(res['draws'].reshape(-1, n), old_idx)
        check_against_oracle(res, res['lik'], doff, seg_off)
        # per-draw LLRs and the partial-sum layout of ldpgta_window_stats
        for p, pair in enumerate(L.LLR_PAIRS):
            col = {'monosomy': 0, 'disomy': 1, 'SPH': 2, 'BPH': 3}
            ref = [O.LLR(l[col[pair[0]]], l[col[pair[1]]]) for l in res['lik']]
            assert np.max(np.abs(res['llr'][p] - np.array(ref))) < 1e-14                # 1-2 ulp + the 2^-53 of the reference's rounded quotient
        for s in range(3):
            lo, hi = doff[seg_off[s]], doff[seg_off[s + 1]]
            _, _, part = eng.window_stats(res['lik'][lo:hi], doff[seg_off[s]:seg_off[s + 1] + 1] - lo, subs)
            assert np.allclose(res['partials'][s], part, rtol=1e-14, atol=1e-13)
        # stats only: same numbers, nothing else downloaded; and repeatable bit for bit
        again = eng.bootstrap_stats(L.HOMOGENEOUS, 99, want=('mean_var', 'stats'))
        assert np.array_equal(again['mean_var'], res['mean_var']) and np.array_equal(again['stats'], res['stats'])
        # replayed draws give the same likelihoods and statistics
        rep = eng.replay_stats(L.HOMOGENEOUS, res['draws'], want=('lik', 'mean_var', 'stats'))
        assert np.array_equal(rep['lik'], res['lik']) and np.array_equal(rep['stats'], res['stats'])
        # bin_stats on the resident LLRs still works after the fused call -- also after a statistics-only call, which does not
        # write the LLR matrix (it is then formed from the resident likelihoods on first use)
        eng.bootstrap_stats(L.HOMOGENEOUS, 99, want=('stats',))
        bins_lazy = eng.bin_stats(None, doff, [[0, 10], [10, 57]])
        eng.bootstrap_stats(L.HOMOGENEOUS, 99, want=('llr',))
        bins = eng.bin_stats(None, doff, [[0, 10], [10, 57]])
        assert bins.shape == (5, 2, 2) and np.isfinite(bins).all()
        assert np.array_equal(bins_lazy, bins)


def test_fused_consecutive_rows_and_two_groups():
    rng = np.random.default_rng(12)
    rows0, woff, _, R = make_case(rng, 33, 7, 30, n_hap=150, shuffle=False)
    rows1 = random_rows(rng, len(rows0), 90)
    doff = np.concatenate([[0], np.cumsum(rng.integers(2, 20, size=len(R)))]).astype(np.int64)
    for kind, prop in ((L.RECENT_ADMIXTURE, None), (L.DISTANT_ADMIXTURE, [0.3, 0.7])):
        with L.Engine([150, 90]) as eng:
            eng.upload_read_masks(0, rows0)
            eng.upload_read_masks(1, rows1)
            old = eng.bootstrap_uniform(kind, 3, woff, np.arange(len(rows0), dtype=np.int32), doff, 5 + 3, proportions=prop)
            eng.set_windows(woff, None, doff, 3)
            res = eng.bootstrap_stats(kind, 5, proportions=prop, want=('lik', 'mean_var', 'stats'))
            assert np.array_equal(res['lik'], old)
            check_against_oracle(res, res['lik'], doff, np.array([0, len(R)]))


def test_fused_mixed_draw_sizes():
    rng = np.random.default_rng(13)
    n_hap = 200
    rows, woff, wreads, R = make_case(rng, 40, 8, 30, n_hap=n_hap)
    sizes = rng.choice([3, 4, 6], size=len(R)).astype(np.int32)
    per = rng.integers(3, 12, size=len(R))
    doff = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    seg_off = np.array([0, 15, 40], dtype=np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        eng.set_windows(woff, wreads, doff, sizes, segment_offsets=seg_off)
        res = eng.bootstrap_stats(L.HOMOGENEOUS, 1234, want=('lik', 'draws', 'mean_var', 'stats'))
        # the unfused route: one ldpgta_bootstrap_uniform call per draw size over all windows
        want = np.empty_like(res['lik'])
        draws = {}
        for k in (3, 4, 6):
            pk = np.where(sizes == k, per, 0)
            dk = np.concatenate([[0], np.cumsum(pk)]).astype(np.int64)
            lik_k, idx_k = eng.bootstrap_uniform(L.HOMOGENEOUS, k, woff, wreads, dk, 1234 + k, want_draws=True)
            at = 0
            for w in np.flatnonzero(sizes == k):
                want[doff[w]:doff[w + 1]] = lik_k[at:at + per[w]]
                draws[w] = idx_k[at:at + per[w]]
                at += per[w]
        assert np.array_equal(res['lik'], want)
        flat = np.concatenate([draws[w].ravel() for w in range(len(R))])
        assert np.array_equal(res['draws'], flat)
        check_against_oracle(res, res['lik'], doff, seg_off)
        # every draw stays inside its window and holds distinct reads
        at = 0
        for w in range(len(R)):
            members = set(wreads[woff[w]:woff[w + 1]].tolist())
            for _ in range(per[w]):
                d = res['draws'][at:at + sizes[w]].tolist()
                assert len(set(d)) == sizes[w] and set(d) <= members
                at += sizes[w]
        rep = eng.replay_stats(L.HOMOGENEOUS, res['draws'], want=('lik', 'stats'))
        assert np.array_equal(rep['lik'], res['lik']) and np.array_equal(rep['stats'], res['stats'])
        with pytest.raises(L.LdpgtaError):
            bad = res['draws'].copy()
            bad[7] = len(rows)
            eng.replay_stats(L.HOMOGENEOUS, bad)


def test_chunked_column_sums_long_segments():
    """ Segments of hundreds of windows with ragged draw counts <= 32 (the chunked reductions: k_window_stats_chunks +
    k_segment_from_chunks, every row of the segment kernel busy, an empty segment, a one-window segment) against the
    reductions of ldpgta_window_stats on the same likelihoods (k_window_stats + k_segment_partials) and the oracle. """
    rng = np.random.default_rng(21)
    n_hap = 96
    rows, woff, wreads, R = make_case(rng, 1500, 4, 9, n_hap=n_hap, shuffle=False)
    per = rng.integers(5, 33, size=len(R))
    per[rng.integers(0, len(R), 40)] = 32
    doff = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    seg_off = np.array([0, 811, 811, 812, 1500], dtype=np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        eng.set_windows(woff, None, doff, 3, segment_offsets=seg_off)
        res = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('lik', 'mean_var', 'partials', 'stats'))
        only = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('partials', 'stats'))
        assert np.array_equal(only['partials'], res['partials']) and np.array_equal(only['stats'], res['stats'], equal_nan=True)
        for s in (0, 2, 3):
            lo, hi = doff[seg_off[s]], doff[seg_off[s + 1]]
            cols = int(per[seg_off[s]:seg_off[s + 1]].min())
            _, mv, part = eng.window_stats(res['lik'][lo:hi], doff[seg_off[s]:seg_off[s + 1] + 1] - lo, cols)
            assert np.array_equal(mv, res['mean_var'][:, seg_off[s]:seg_off[s + 1]], equal_nan=True)
            assert np.allclose(res['partials'][s][:, :cols + 3], part, rtol=1e-13, atol=1e-12)
            assert not res['partials'][s][:, cols + 3:].any()
        keep = np.array([0, 300], dtype=np.int64)            # the oracle's statistics() on the first 300 windows as one segment
        eng.set_windows(woff[:301], None, doff[:301], 3, segment_offsets=keep)
        sub = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('lik', 'mean_var', 'stats'))
        check_against_oracle(sub, sub['lik'], doff[:301], keep)


def test_fused_many_draws_per_window():
    """ 1000 subsamples per window (configs[4]); one window short of the others truncates the columns (:73). """
    rng = np.random.default_rng(14)
    n_hap = 128
    rows, woff, wreads, R = make_case(rng, 12, 12, 40, n_hap=n_hap)
    per = np.full(len(R), 1000)
    per[5] = 377
    doff = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        info = eng.set_windows(woff, wreads, doff, 4)
        assert info['ncp'] == 377
        res = eng.bootstrap_stats(L.HOMOGENEOUS, 7, want=('lik', 'mean_var', 'partials', 'stats'))
        check_against_oracle(res, res['lik'], doff, np.array([0, len(R)]))


def test_window_resident_kernel_many_draws_per_window():
    """ Plans with hundreds of draws per window, 4 reads per draw, 5008 haplotypes and windows that are runs of consecutive
    rows take k_small_win (a window's masks resident in shared memory): likelihoods bit-identical to the per-draw gather
    kernel behind ldpgta_bootstrap_uniform, statistics equal to the oracle's, ragged windows (4 reads .. 170 reads, draw
    counts not multiples of the tile or the batch). """
    rng = np.random.default_rng(29)
    n_hap = 5008
    R = np.array([4, 170, 33, 5, 96, 150, 61, 12, 140])
    rows = random_rows(rng, int(R.sum()), n_hap)
    woff = np.concatenate([[0], np.cumsum(R)]).astype(np.int64)
    per = np.array([300, 1000, 257, 999, 512, 1001, 333, 700, 64])
    doff = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    seg_off = np.array([0, 4, 9], dtype=np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        old_lik, old_idx = eng.bootstrap_uniform(L.HOMOGENEOUS, 4, woff, np.arange(len(rows), dtype=np.int32), doff, 5 + 4, want_draws=True)
        launches = eng.launch_count
        eng.set_windows(woff, None, doff, 4, segment_offsets=seg_off)
        res = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('lik', 'draws', 'mean_var', 'stats'))
        assert np.array_equal(res['draws'].reshape(-1, 4), old_idx)
        assert np.array_equal(res['lik'], old_lik)
        check_against_oracle(res, res['lik'], doff, seg_off)
        again = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('lik', 'stats'))
        assert np.array_equal(again['lik'], res['lik']) and np.array_equal(again['stats'], res['stats'])
        rep = eng.replay_stats(L.HOMOGENEOUS, res['draws'], want=('lik', 'stats'))
        assert np.array_equal(rep['lik'], res['lik']) and np.array_equal(rep['stats'], res['stats'])
        del launches


def test_window_resident_kernel_on_read_lists():
    """ The same kernel for windows given as READ LISTS (what packed.py hands over: reads in arbitrary row order, rows that no
    window uses, rows shared by neighbouring windows): the plan keeps a window-ordered copy of the masks and the sampler's
    positions inside the windows.  Draws and likelihoods bit-identical to the per-draw gather kernel, statistics equal to the
    oracle's; the copy is made once (one more launch in the first call than in the second) and again after the read masks
    change; draws of the caller (replay) give the same numbers through k_small. """
    rng = np.random.default_rng(31)
    n_hap = 5008
    R = np.array([4, 170, 33, 5, 96, 150, 61, 12, 140])
    n_rows = int(R.sum()) + 50                                     # 50 rows belong to no window
    rows = random_rows(rng, n_rows, n_hap)
    wreads = rng.permutation(n_rows)[:int(R.sum())].astype(np.int32)
    woff = np.concatenate([[0], np.cumsum(R)]).astype(np.int64)
    wreads[woff[2]:woff[2] + 3] = wreads[woff[1]:woff[1] + 3]      # neighbours share reads
    per = np.array([300, 1000, 257, 999, 512, 1001, 333, 700, 64])
    doff = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    seg_off = np.array([0, 4, 9], dtype=np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        old_lik, old_idx = eng.bootstrap_uniform(L.HOMOGENEOUS, 4, woff, wreads, doff, 5 + 4, want_draws=True)
        eng.set_windows(woff, wreads, doff, 4, segment_offsets=seg_off)
        before = eng.launch_count
        res = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('lik', 'draws', 'mean_var', 'stats'))
        first = eng.launch_count - before
        assert np.array_equal(res['draws'].reshape(-1, 4), old_idx)
        assert np.array_equal(res['lik'], old_lik)
        check_against_oracle(res, res['lik'], doff, seg_off)
        before = eng.launch_count
        again = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('lik', 'draws', 'mean_var', 'stats'))
        assert first - (eng.launch_count - before) == 1            # the window-ordered copy was gathered in the first call only
        assert np.array_equal(again['lik'], res['lik']) and np.array_equal(again['stats'], res['stats'])
        rep = eng.replay_stats(L.HOMOGENEOUS, res['draws'], want=('lik', 'stats'))
        assert np.array_equal(rep['lik'], res['lik']) and np.array_equal(rep['stats'], res['stats'])
        # statistics only, another seed: equal to the per-draw kernel on the same draws
        other = eng.bootstrap_stats(L.HOMOGENEOUS, 6, want=('draws', 'stats', 'mean_var'))
        rep = eng.replay_stats(L.HOMOGENEOUS, other['draws'], want=('stats', 'mean_var'))
        assert np.array_equal(rep['stats'], other['stats']) and np.array_equal(rep['mean_var'], other['mean_var'], equal_nan=True)
        # nothing but statistics asked for: the sampler writes positions only (no rows of the table) -- same numbers, and the
        # rows are back when a later call wants them
        only = eng.bootstrap_stats(L.HOMOGENEOUS, 6, want=('stats', 'mean_var'))
        assert np.array_equal(only['stats'], other['stats']) and np.array_equal(only['mean_var'], other['mean_var'], equal_nan=True)
        back = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('draws', 'stats'))
        assert np.array_equal(back['draws'], res['draws']) and np.array_equal(back['stats'], res['stats'])
        # new read masks under the same plan: the copy follows
        rows2 = random_rows(rng, n_rows, n_hap, density=0.4)
        eng.upload_read_masks(0, rows2)
        new_lik, new_idx = eng.bootstrap_uniform(L.HOMOGENEOUS, 4, woff, wreads, doff, 5 + 4, want_draws=True)
        eng.set_windows(woff, wreads, doff, 4, segment_offsets=seg_off)
        res2 = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('lik',))
        assert np.array_equal(res2['lik'], new_lik) and not np.array_equal(new_lik, old_lik)
        eng.upload_read_masks(0, rows)                             # ... also when the plan stays
        res3 = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('lik',))
        assert np.array_equal(res3['lik'], old_lik)
        # a table with fewer rows than the windows were checked against: refused, not read out of bounds
        eng.upload_read_masks(0, rows[:100])
        with pytest.raises(L.LdpgtaError):
            eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('stats',))


def test_sliced_statistics_download_into_pinned_buffers():
    """ With page-locked result buffers and enough chunks the window statistics go down in slices under the reductions of
    the later windows (strided 2-D copies into mean_var[5][n_windows][2]); same numbers as the one-copy route that
    pageable buffers take, also when the likelihoods are downloaded in the same call. """
    rng = np.random.default_rng(23)
    n_hap, n_win = 64, 9001                                       # 1126 chunks: two slices
    rows, woff, wreads, R = make_case(rng, n_win, 3, 6, n_hap=n_hap, shuffle=False)
    per = rng.integers(2, 6, size=n_win)
    doff = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    seg_off = np.array([0, 4000, n_win], dtype=np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        info = eng.set_windows(woff, None, doff, 2, segment_offsets=seg_off)
        plain = eng.bootstrap_stats(L.HOMOGENEOUS, 77, want=('lik', 'mean_var', 'partials', 'stats'))
        pins = {'mean_var': L.PinnedArray((5, n_win, 2), np.float64), 'stats': L.PinnedArray((2, 5, 3), np.float64),
                'lik': L.PinnedArray((int(doff[-1]), 4), np.float64)}
        for p_ in pins.values():
            p_.array[...] = -7.0
        out = {'mean_var': pins['mean_var'].array, 'stats': pins['stats'].array}
        eng.bootstrap_stats(L.HOMOGENEOUS, 77, out=out)
        assert np.array_equal(out['mean_var'], plain['mean_var'], equal_nan=True) and np.array_equal(out['stats'], plain['stats'])
        out['mean_var'][...] = -7.0
        out['lik'] = pins['lik'].array
        eng.bootstrap_stats(L.HOMOGENEOUS, 77, out=out)
        assert np.array_equal(out['lik'], plain['lik']) and np.array_equal(out['mean_var'], plain['mean_var'], equal_nan=True)
        rep = eng.replay_stats(L.HOMOGENEOUS, eng.bootstrap_stats(L.HOMOGENEOUS, 77, want=('draws',))['draws'],
                               out={'mean_var': pins['mean_var'].array, 'stats': pins['stats'].array})
        assert np.array_equal(rep['mean_var'], plain['mean_var'], equal_nan=True) and np.array_equal(rep['stats'], plain['stats'])
    for p_ in pins.values():
        p_.free()


@pytest.mark.parametrize('most', [33, 256, 300, 512, 2048, 2100])
def test_fused_window_sizes_of_the_block_kernels(most):
    """ Windows of 33..2048 draws go through k_window_stats_block<1|2|4|8> (a block per window, column sums per chunk of
    windows), more than 2048 through k_draw_llr + k_window_stats + k_segment_partials; ragged windows, two segments, one
    of them longer than a chunk. """
    rng = np.random.default_rng(100 + most)
    n_hap = 70
    rows, woff, wreads, R = make_case(rng, 21, 5, 12, n_hap=n_hap)
    per = rng.integers(max(2, most // 3), most + 1, size=len(R))
    per[3] = most
    doff = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    seg_off = np.array([0, 4, 21], dtype=np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        info = eng.set_windows(woff, wreads, doff, 2, segment_offsets=seg_off)
        assert info['ncp'] == max(per[:4].min(), per[4:].min())
        res = eng.bootstrap_stats(L.HOMOGENEOUS, 3, want=('lik', 'llr', 'mean_var', 'partials', 'stats'))
        only = eng.bootstrap_stats(L.HOMOGENEOUS, 3, want=('mean_var', 'partials', 'stats'))
        for k in ('mean_var', 'partials', 'stats'):
            assert np.array_equal(only[k], res[k], equal_nan=True), k
        lazy = eng.bin_stats(None, doff, [[0, 4], [4, 21]])
        assert np.array_equal(lazy, eng.bin_stats(res['llr'], doff, [[0, 4], [4, 21]]), equal_nan=True)
        check_against_oracle(res, res['lik'], doff, seg_off)
        for s in range(2):
            lo, hi = doff[seg_off[s]], doff[seg_off[s + 1]]
            cols = int(per[seg_off[s]:seg_off[s + 1]].min())
            _, mv, part = eng.window_stats(res['lik'][lo:hi], doff[seg_off[s]:seg_off[s + 1] + 1] - lo, cols)
            assert np.allclose(res['partials'][s][:, :cols + 3], part, rtol=1e-13, atol=1e-12)
            assert not res['partials'][s][:, cols + 3:].any()


def test_shards_of_a_chromosome_add_up():
    """ Two contexts own halves of the windows of two chromosomes (segment_cols / segment_first taken from the whole
    chromosome, as ranks of a multi-GPU run do); their segment partials add up to the single-context result. """
    rng = np.random.default_rng(15)
    n_hap = 96
    rows, woff, wreads, R = make_case(rng, 30, 8, 25, n_hap=n_hap, shuffle=False)
    per = rng.integers(4, 9, size=len(R))
    doff = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    seg_off = np.array([0, 13, 30], dtype=np.int64)
    cols = np.array([per[:13].min(), per[13:].min()], dtype=np.int32)
    first = np.array([per[0], per[13]], dtype=np.int32)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        eng.set_windows(woff, None, doff, 4, segment_offsets=seg_off)
        whole = eng.bootstrap_stats(L.HOMOGENEOUS, 3, want=('lik', 'draws', 'partials', 'stats'))
        draws = whole['draws'].reshape(-1, 4)
        total = np.zeros_like(whole['partials'])
        for lo, hi in ((0, 9), (9, 30)):                         # the cut falls inside chromosome 0
            so = np.clip(seg_off, lo, hi) - lo
            eng.set_windows(woff[lo:hi + 1], None, doff[lo:hi + 1] - doff[lo], 4, segment_offsets=so, segment_cols=cols, segment_first=first)
            total += eng.replay_stats(L.HOMOGENEOUS, draws[doff[lo]:doff[hi]], want=('partials',))['partials']
        assert np.allclose(total, whole['partials'], rtol=1e-13, atol=1e-12)
        got = np.stack([L.finish_chromosome(np.ascontiguousarray(total[s][:, :cols[s] + 3]), int(cols[s]), int(first[s])) for s in range(2)])
        assert np.max(np.abs(got - whole['stats'])) < 1e-12


def test_plan_errors():
    rng = np.random.default_rng(16)
    rows = random_rows(rng, 20, 64)
    with L.Engine([64]) as eng:
        with pytest.raises(L.LdpgtaError):
            eng.bootstrap_stats(L.HOMOGENEOUS, 1)                              # no plan yet
        eng.upload_read_masks(0, rows)
        with pytest.raises(L.LdpgtaError):
            eng.set_windows([0, 3, 20], None, [0, 4, 8], 4)                    # window 0 has 3 reads < 4
        with pytest.raises(L.LdpgtaError):
            eng.set_windows([0, 10, 20], None, [0, 4, 4], 4)                   # window 1 has no draws
        with pytest.raises(L.LdpgtaError):
            eng.set_windows([0, 10, 20], [0] * 19 + [20], [0, 4, 8], 4)        # read row 20 does not exist
        with pytest.raises(L.LdpgtaError):
            eng.set_windows([0, 10, 21], None, [0, 4, 8], 4)                   # rows beyond the read-mask table
        eng.set_windows(None, None, [0, 4, 8], 4)                              # replay-only plan
        with pytest.raises(L.LdpgtaError):
            eng.bootstrap_stats(L.HOMOGENEOUS, 1)
        out = eng.replay_stats(L.HOMOGENEOUS, rng.integers(0, 20, size=(8, 4)), want=('lik', 'stats'))
        assert out['lik'].shape == (8, 4)
