"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI
(ctypes) and through the reference-shaped Python classes, against the oracle and the golden
vectors produced by the unmodified reference.

Bars (BASELINE.json north_star): joint-frequency popcounts bit-exact; per-draw likelihoods
within 1e-12 relative; chromosome-level LLR mean / std within 1e-9.
"""
import random

import numpy as np
import pytest

import coracle as C
import ldpgta_oracle as O
from conftest import CHROM_CONFIGS, chrom_case, rel_err

import ldpgta_b200 as L
from ldpgta_b200 import driver

pytestmark = pytest.mark.gpu

LIK_TOL = 1e-12          # per-draw likelihoods, relative
STAT_TOL = 1e-9          # chromosome-level LLR mean / std


def max_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0


def ld_rows(rng, n_rows, n_hap, founders=10, density=0.6):
    """ Correlated rows (mosaic of a few founder patterns) so deep intersections stay non-empty. """
    words = (n_hap + 63) // 64
    base = rng.integers(0, 2 ** 64, size=(founders, words), dtype=np.uint64)
    extra = rng.integers(0, 2 ** 64, size=(n_rows, words), dtype=np.uint64)
    pick = rng.integers(0, founders, size=n_rows)
    rows = base[pick] | (extra & rng.integers(0, 2 ** 64, size=(n_rows, words), dtype=np.uint64))
    if density > 0.7:
        rows |= rng.integers(0, 2 ** 64, size=(n_rows, words), dtype=np.uint64)
    tail = n_hap - 64 * (words - 1)
    if tail < 64:
        rows[:, -1] &= np.uint64((1 << tail) - 1)
    return rows


def uniform_draws(rng, n_reads, n, n_draws):
    idx = np.stack([rng.choice(n_reads, size=n, replace=False) for _ in range(n_draws)]).astype(np.int32)
    return idx.ravel(), np.arange(0, n * n_draws + 1, n, dtype=np.int64)


# ----------------------------------------------------------------------------------------
# golden vectors of the reference, through the reference-shaped classes
# ----------------------------------------------------------------------------------------

def test_testmodule_golden_three_model_classes(golden_testmodule):
    g = golden_testmodule
    obs, leg, hap, nh = g['obs_tab'], g['leg_tab'], g['hap_tab'], g['number_of_haplotypes']
    models = L.ImplicitModels()
    homog = {grp: L.homogeneous(obs, leg, {grp: hap[grp]}, {grp: nh[grp]}, models) for grp in hap}
    recent = L.recent_admixture(obs, leg, hap, nh, models)
    distant = [L.distant_admixture(obs, leg, hap, nh, models, mk) for mk in g['makeups']]
    worst = 0.0
    for grp, m in homog.items():
        assert m.fraction_of_matches == g['fraction_of_matches'][grp]
    for rec in g['records']:
        al = rec['alleles']
        for grp, m in homog.items():
            assert m.joint_frequencies_combo(al, normalize=False) == rec['homog_counts_' + grp]      # bit-exact
            worst = max(worst, max_rel(m.likelihoods(al), rec['homog_general_' + grp]))
            worst = max(worst, max_rel(m.get_likelihoods(al), rec['homog_get_' + grp]))
        worst = max(worst, max_rel(recent.likelihoods(al), rec['recent_general']))
        worst = max(worst, max_rel(recent.get_likelihoods(al), rec['recent_get']))
        for grp in hap:
            assert recent.joint_frequencies_combo(al, grp, normalize=False) == rec['homog_counts_' + grp]
        for i, d in enumerate(distant):
            freq = d.joint_frequencies_combo(al)
            assert max(rel_err(freq[s], v) for s, v in rec['distant%d_freq' % i].items()) < 1e-15
            worst = max(worst, max_rel(d.likelihoods(al), rec['distant%d_general' % i]))
            if rec['n'] == 5:
                with pytest.raises(AttributeError):
                    d.get_likelihoods(al)
            else:
                worst = max(worst, max_rel(d.get_likelihoods(al), rec['distant%d_get' % i]))
    assert worst < LIK_TOL, worst
    # closed forms of the single-population model follow CPython's operation order: expect bit equality
    exact = sum(tuple(homog['group0'].get_likelihoods(r['alleles'])) == tuple(r['homog_get_group0'])
                for r in g['records'] if r['n'] in (2, 3, 4))
    total = sum(1 for r in g['records'] if r['n'] in (2, 3, 4))
    assert exact == total, (exact, total)
    with pytest.raises(KeyError):
        homog['group0'].get_likelihoods(((1, 'A'), (2, 'C')))          # unknown allele, like the reference
    for m in list(homog.values()) + [recent] + distant:
        m.close()


@pytest.mark.parametrize('name', CHROM_CONFIGS)
def test_bootstrap_replay_matches_reference(name, golden_inputs):
    args, out = chrom_case(name, golden_inputs)
    cfg = args['cfg']
    lik, windows, matches = L.bootstrap(args['obs_tab'], args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'],
                                        args['ancestral_makeup'], L.ImplicitModels(), cfg['window_size'], cfg['subsamples'],
                                        cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'], cfg['min_HF'],
                                        replay=out['draws'])
    assert list(windows) == list(out['windows'])
    assert matches == out['fraction_of_matches']
    assert list(lik) == list(out['likelihoods'])
    worst = 0.0
    for w, ls in lik.items():
        assert len(ls) == len(out['likelihoods'][w])
        for mine, ref in zip(ls, out['likelihoods'][w]):
            assert mine._fields == ('monosomy', 'disomy', 'SPH', 'BPH')
            worst = max(worst, max_rel(mine, ref))
    assert worst < LIK_TOL, worst
    # statistics() on the device vs the reference's numbers (its likelihoods as input)
    ref_lik = {w: tuple(L.likelihoods_tuple(*l) for l in ls) for w, ls in out['likelihoods'].items()}
    st = L.statistics(ref_lik, out['windows'])
    gs = out['statistics']
    for key in ('num_of_windows', 'reads_mean', 'reads_std', 'window_size_mean', 'window_size_std'):
        assert st[key] == gs[key]
    for pair in L.LLR_PAIRS:
        for key in ('mean_of_mean', 'std_of_mean', 'fraction_of_negative_LLRs'):
            a, b = st['LLRs_per_chromosome'][pair][key], gs['LLRs_per_chromosome'][pair][key]
            assert abs(a - b) <= STAT_TOL * max(1.0, abs(b)), (pair, key, a, b)
        for w, (m, v) in gs['LLRs_per_genomic_window'][pair].items():
            mm, vv = st['LLRs_per_genomic_window'][pair][w]
            assert abs(mm - m) <= STAT_TOL * max(1.0, abs(m)) and abs(vv - v) <= STAT_TOL * max(1.0, abs(v))
    # and end to end: device likelihoods -> device statistics
    st2 = L.statistics(lik, windows)
    for key in ('mean_of_mean', 'std_of_mean'):
        a, b = st2['LLRs_per_chromosome'][('BPH', 'SPH')][key], gs['LLRs_per_chromosome'][('BPH', 'SPH')][key]
        assert abs(a - b) <= STAT_TOL * max(1.0, abs(b))


# ----------------------------------------------------------------------------------------
# C-ABI against the C oracle on seeded inputs
# ----------------------------------------------------------------------------------------

def test_every_small_kernel_tiling():
    """ One panel size per (lanes per draw, words per lane) instance of k_small, plus the unpipelined kernel. """
    rng = np.random.default_rng(4)
    for n_hap in (900, 1500, 2504, 3000, 5008, 6000, 9000):
        rows = ld_rows(rng, 64, n_hap)
        with L.Engine([n_hap]) as eng:
            eng.upload_read_masks(0, rows)
            for n in (2, 3, 4):
                idx, off = uniform_draws(rng, 64, n, 61)              # 61: ragged last tile
                got = eng.likelihoods_batch(L.HOMOGENEOUS, idx, off)
                want = C.batch([rows], [n_hap], idx, off, 0)
                assert max_rel(got, want) < LIK_TOL, (n_hap, n)
                assert np.array_equal(eng.joint_counts(idx[:n])[0], C.joint_counts(rows[idx[:n]]))


@pytest.mark.parametrize('n_hap', [50, 64, 65, 331, 2504, 5008])
def test_homogeneous_all_draw_sizes(n_hap):
    rng = np.random.default_rng(n_hap)
    rows = ld_rows(rng, 96, n_hap)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        for n in range(2, 19):
            nd = 40 if n <= 8 else (6 if n <= 13 else (2 if n <= 16 else 1))
            idx, off = uniform_draws(rng, 96, n, nd)
            got = eng.likelihoods_batch(L.HOMOGENEOUS, idx, off)
            want = C.batch([rows], [n_hap], idx, off, 0)
            assert max_rel(got, want) < LIK_TOL, (n, max_rel(got, want))
            cnt = eng.joint_counts(idx[:n])
            assert np.array_equal(cnt[0], C.joint_counts(rows[idx[:n]])), n          # bit-exact popcounts


def test_dense_panel_deep_intersections_n16():
    """ Dense correlated rows: the 16-read intersections are non-zero, so every term of the
    three-block sum contributes. """
    rng = np.random.default_rng(16)
    n_hap = 5008
    rows = ld_rows(rng, 32, n_hap, founders=3, density=0.9)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        for n in (12, 15, 16):
            idx, off = uniform_draws(rng, 32, n, 3)
            cnt = eng.joint_counts(idx[:n])
            ref = C.joint_counts(rows[idx[:n]])
            assert ref[-1] > 0
            assert np.array_equal(cnt[0], ref)
            got = eng.likelihoods_batch(L.HOMOGENEOUS, idx, off)
            want = C.batch([rows], [n_hap], idx, off, 0)
            assert max_rel(got, want) < LIK_TOL, (n, max_rel(got, want))


@pytest.mark.parametrize('sizes', [(331, 200), (2504, 2504), (50, 5008)])
def test_recent_and_distant_admixture(sizes):
    rng = np.random.default_rng(sum(sizes))
    rows = [ld_rows(rng, 80, s) for s in sizes]
    with L.Engine(list(sizes)) as eng:
        for g in range(2):
            eng.upload_read_masks(g, rows[g])
        for n in range(2, 19):                                  # every draw size the device accepts, both two-group models
            nd = 24 if n <= 8 else (3 if n <= 14 else 1)
            idx, off = uniform_draws(rng, 80, n, nd)
            got = eng.likelihoods_batch(L.RECENT_ADMIXTURE, idx, off)
            want = C.batch(_common_width(rows), list(sizes), idx, off, 1)
            assert max_rel(got, want) < LIK_TOL, ('recent', n, max_rel(got, want))
            if n != 5:
                prop = [0.3, 0.7]
                got = eng.likelihoods_batch(L.DISTANT_ADMIXTURE, idx, off, proportions=prop)
                want = C.batch(_common_width(rows), list(sizes), idx, off, 2, prop)
                assert max_rel(got, want) < LIK_TOL, ('distant', n, max_rel(got, want))
            cnt = eng.joint_counts(idx[:n])
            for g in range(2):
                assert np.array_equal(cnt[g], C.joint_counts(rows[g][idx[:n]]))


@pytest.mark.parametrize('sizes', [(11585, 11585), (12000, 9000), (16383, 16383), (16400, 300)])
def test_recent_admixture_accumulator_width(sizes):
    """ Groups around the sizes at which the polynomial's 32-bit convolution sums would overflow (recent admixture adds two
    convolutions of 16 products: 32 * N0 * N1 < 2^32 holds up to 11585 haplotypes per group), with dense rows so that the
    counts really are that large. """
    rng = np.random.default_rng(sizes[0])
    rows = [ld_rows(rng, 48, s, founders=2, density=0.95) | np.uint64(0x7fffffffffffffff) for s in sizes]
    for r, s in zip(rows, sizes):
        tail = s - 64 * ((s + 63) // 64 - 1)
        if tail < 64:
            r[:, -1] &= np.uint64((1 << tail) - 1)
    with L.Engine(list(sizes)) as eng:
        for g in range(2):
            eng.upload_read_masks(g, rows[g])
        for n in (5, 6, 9, 12, 13):
            idx, off = uniform_draws(rng, 48, n, 4 if n < 12 else 2)
            got = eng.likelihoods_batch(L.RECENT_ADMIXTURE, idx, off)
            want = C.batch(_common_width(rows), list(sizes), idx, off, 1)
            assert max_rel(got, want) < LIK_TOL, (sizes, n, max_rel(got, want))
            if n != 5:
                # distant admixture of two groups runs the same integer convolutions (poly_general.cuh: Cubic2)
                prop = [0.35, 0.65]
                got = eng.likelihoods_batch(L.DISTANT_ADMIXTURE, idx, off, proportions=prop)
                want = C.batch(_common_width(rows), list(sizes), idx, off, 2, prop)
                assert max_rel(got, want) < LIK_TOL, ('distant', sizes, n, max_rel(got, want))


@pytest.mark.parametrize('model,sizes', [('homogeneous', (5792,)), ('homogeneous', (5793,)), ('homogeneous', (8191,)),
                                         ('homogeneous', (8192,)), ('homogeneous', (11585,)), ('homogeneous', (11586,)),
                                         ('recent', (4096, 4096)), ('recent', (4097, 3000)), ('recent', (5792, 5792)),
                                         ('recent', (5793, 100)), ('recent', (8191, 8191)), ('recent', (8192, 8192))])
def test_polynomial_run_lengths(model, sizes):
    """ From 15 reads the integer polynomials sum the convolutions of up to 2, 4 or 8 colourings that share the X row in 32-bit
    accumulators before the dot product (poly_general.cuh); how many depends on the panel size.  Panels on both sides of every
    threshold, with nearly all-ones rows so that the sums come as close to 2^32 as they can. """
    rng = np.random.default_rng(sum(sizes))
    rows = [ld_rows(rng, 24, s, founders=2, density=0.95) | np.uint64(0x7fffffffffffffff) for s in sizes]
    for r, s in zip(rows, sizes):
        tail = s - 64 * ((s + 63) // 64 - 1)
        if tail < 64:
            r[:, -1] &= np.uint64((1 << tail) - 1)
    kind, code = (L.HOMOGENEOUS, 0) if model == 'homogeneous' else (L.RECENT_ADMIXTURE, 1)
    with L.Engine(list(sizes)) as eng:
        for g in range(len(sizes)):
            eng.upload_read_masks(g, rows[g])
        for n in (15, 16):
            idx, off = uniform_draws(rng, 24, n, 2)
            got = eng.likelihoods_batch(kind, idx, off)
            want = C.batch(_common_width(rows), list(sizes), idx, off, code)
            assert max_rel(got, want) < LIK_TOL, (model, sizes, n, max_rel(got, want))
            if model == 'recent':                                 # distant admixture of two groups shares the run lists
                prop = [0.8, 0.2]
                got = eng.likelihoods_batch(L.DISTANT_ADMIXTURE, idx, off, proportions=prop)
                want = C.batch(_common_width(rows), list(sizes), idx, off, 2, prop)
                assert max_rel(got, want) < LIK_TOL, ('distant', sizes, n, max_rel(got, want))


@pytest.mark.parametrize('sizes', [(2504, 2504), (2100, 2900), (3000, 2008), (2048, 2560)])
def test_two_groups_in_one_row(sizes):
    """ Batches of at least 4096 four-read draws of a two-group panel whose rows together fill the 80-word tile run on the
    combined-row kernel (k_small_both: one gather per read for both groups, the boundary inside the third slot row); the
    sizes put the boundary at different lanes, including none (2048 = 32 words exactly). """
    rng = np.random.default_rng(sizes[0])
    rows = [ld_rows(rng, 300, s) for s in sizes]
    idx, off = uniform_draws(rng, 300, 4, 6000)
    with L.Engine(list(sizes)) as eng:
        for g in range(2):
            eng.upload_read_masks(g, rows[g])
        got = eng.likelihoods_batch(L.RECENT_ADMIXTURE, idx, off)
        want = C.batch(_common_width(rows), list(sizes), idx, off, 1)
        assert max_rel(got, want) < LIK_TOL, ('recent', sizes, max_rel(got, want))
        prop = [0.25, 0.75]
        got = eng.likelihoods_batch(L.DISTANT_ADMIXTURE, idx, off, proportions=prop)
        want = C.batch(_common_width(rows), list(sizes), idx, off, 2, prop)
        assert max_rel(got, want) < LIK_TOL, ('distant', sizes, max_rel(got, want))
        # new masks for one group: the combined rows are rebuilt
        rows[1] = ld_rows(rng, 300, sizes[1])
        eng.upload_read_masks(1, rows[1])
        got = eng.likelihoods_batch(L.RECENT_ADMIXTURE, idx, off)
        want = C.batch(_common_width(rows), list(sizes), idx, off, 1)
        assert max_rel(got, want) < LIK_TOL, ('recent, new masks', sizes, max_rel(got, want))


def _common_width(rows):
    W = max(r.shape[1] for r in rows)
    out = []
    for r in rows:
        p = np.zeros((r.shape[0], W), dtype=np.uint64)
        p[:, :r.shape[1]] = r
        out.append(p)
    return out


def test_three_group_distant_admixture():
    rng = np.random.default_rng(3)
    sizes = [120, 331, 77]
    rows = [ld_rows(rng, 40, s) for s in sizes]
    prop = [0.5, 0.3, 0.2]
    with L.Engine(sizes) as eng:
        for g in range(3):
            eng.upload_read_masks(g, rows[g])
        for n in (2, 3, 4, 6, 9, 12, 16):
            idx, off = uniform_draws(rng, 40, n, 10 if n <= 9 else (3 if n <= 12 else 1))
            got = eng.likelihoods_batch(L.DISTANT_ADMIXTURE, idx, off, proportions=prop)
            want = C.batch(_common_width(rows), sizes, idx, off, 2, prop)
            assert max_rel(got, want) < LIK_TOL


def test_large_draws_vs_reference_build(golden_inputs):
    """ 10..16 reads per draw (17, 18 when generated) through the three model classes against the UNMODIFIED reference run
    with MAKE_STATISTICAL_MODEL.BUILD tables (tests/golden/make_golden_large.py): joint-frequency tables bit-exact,
    likelihoods <= 1e-12 -- no oracle in between. """
    import hashlib
    from conftest import load_golden
    g = load_golden('large_draws.p.bz2')
    inp = golden_inputs
    obs, leg, hap, nh = inp['obs_tabs']['disomy'], inp['leg_tab'], inp['hap_tab'], inp['number_of_haplotypes']
    models = L.ImplicitModels()
    homog = L.homogeneous(obs, leg, {'G0': hap['G0']}, {'G0': nh['G0']}, models)
    recent = L.recent_admixture(obs, leg, hap, nh, models)
    worst = 0.0
    try:
        for n, rec in sorted(g['records'].items()):
            distant = L.distant_admixture(obs, leg, hap, nh, models, rec['makeup'])
            try:
                al = rec['alleles']
                counts = homog.joint_frequencies_combo(al, normalize=False)
                arr = np.zeros(1 << n, dtype='<u4')
                for sub, c in counts.items():
                    arr[sub] = c
                assert hashlib.sha256(arr.tobytes()).hexdigest() == rec['counts_sha256_G0'], n
                for model, key in ((homog, 'homog_G0'), (recent, 'recent'), (distant, 'distant')):
                    got = model.likelihoods(al)
                    err = max(rel_err(a, b) for a, b in zip(got, rec[key]))
                    assert err < LIK_TOL, (n, key, err)
                    worst = max(worst, err)
            finally:
                distant.close()
    finally:
        homog.close(); recent.close()
    assert set(range(10, 19)) <= set(g['records'])


def test_chr21_full_panel_vs_reference_bootstrap():
    """ configs[0] of BASELINE.json at full panel size (5008 haplotypes): the recorded random.sample stream of the
    UNMODIFIED reference bootstrap() replayed through the drop-in bootstrap() and through the device-resident packed
    pipeline; likelihoods <= 1e-12, window and chromosome statistics <= 1e-9 against the reference's statistics(). """
    import sys
    from conftest import GOLDEN, load_golden
    sys.path.insert(0, GOLDEN)
    import make_golden_large as G
    from ldpgta_b200.packed import PackedSample
    out = load_golden('chr21_5008.p.bz2')
    cfg, gold = out['cfg'], out['statistics']
    leg, rows = G.chr21_panel()
    hap, nh = {'ALL': rows}, {'ALL': G.CHR21_HAP}
    lik, windows, matches = driver.bootstrap(out['obs_tab'], leg, hap, nh, {'ALL'}, L.ImplicitModels(), cfg['window_size'], cfg['subsamples'],
                                             cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'], cfg['min_HF'], replay=out['draws'])
    assert list(windows) == list(out['windows']) and matches == out['fraction_of_matches'] and list(lik) == list(out['likelihoods'])
    worst = max(rel_err(a, b) for w in lik for mine, ref in zip(lik[w], out['likelihoods'][w]) for a, b in zip(mine, ref))
    assert worst < LIK_TOL, worst
    st = driver.statistics(lik, windows)
    for pair in L.LLR_PAIRS:
        for key, val in gold['LLRs_per_chromosome'][pair].items():
            assert abs(st['LLRs_per_chromosome'][pair][key] - val) <= STAT_TOL * max(1e-3, abs(val)), (pair, key)
    # the array route: same windows, the same draws as read rows, statistics straight from the device
    panel = L.ResidentPanel(leg, hap, nh)
    try:
        with PackedSample(out['obs_tab'], panel, {'ALL'}, cfg['min_HF']) as ps:
            row_of = {rid: r for r, rid in enumerate(ps.read_names)}
            flat = [row_of[r] for w in out['likelihoods'] for d in out['draws'][w] for r in d]
            res = ps.run(cfg['window_size'], cfg['subsamples'], cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'], replay=flat)
            assert [tuple(b) for b in res.borders.tolist()] == list(out['likelihoods'])
            ref_lik = np.array([l for w in out['likelihoods'] for l in out['likelihoods'][w]], dtype=np.float64)
            assert max_rel(res.likelihoods, ref_lik) < LIK_TOL
            for p_i, pair in enumerate(L.LLR_PAIRS):
                for k, w in enumerate(out['likelihoods']):
                    m, v = gold['LLRs_per_genomic_window'][pair][w]
                    gm, gv = res.window_mean_var[p_i, k]
                    assert abs(gm - m) <= STAT_TOL * max(1e-3, abs(m)) and abs(gv - v) <= STAT_TOL * max(1e-3, abs(v))
                for key, val in gold['LLRs_per_chromosome'][pair].items():
                    assert abs(res.llr_per_chromosome()[pair][key] - val) <= STAT_TOL * max(1e-3, abs(val)), (pair, key)
    finally:
        panel.close()


def test_ragged_batch_and_scatter_order():
    """ Draw sizes mixed inside one batch (windows with R <= max_reads draw R-1 reads). """
    rng = np.random.default_rng(11)
    n_hap = 5008
    rows = ld_rows(rng, 200, n_hap)
    sizes = rng.choice([2, 3, 4, 4, 4, 5, 6, 8, 9, 12], size=300)
    idx = np.concatenate([rng.choice(200, size=s, replace=False) for s in sizes]).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        got = eng.likelihoods_batch(L.HOMOGENEOUS, idx, off)
    want = C.batch([rows], [n_hap], idx, off, 0)
    assert max_rel(got, want) < LIK_TOL


def test_device_side_mask_construction():
    """ ldpgta_upload_allele_rows (complement of REF alleles) + ldpgta_build_read_masks. """
    rng = random.Random(9)
    n_hap = 331
    panel = [rng.getrandbits(n_hap) for _ in range(300)]
    comp = [rng.random() < 0.5 for _ in panel]
    ones = (1 << n_hap) - 1
    reads = [[rng.randrange(300) for _ in range(rng.choice((1, 1, 2, 3, 5)))] for _ in range(120)]
    masks = []
    for r in reads:
        m = ones
        for a in r:
            m &= panel[a] ^ ones if comp[a] else panel[a]
        masks.append(m)
    off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    flat = np.array([a for r in reads for a in r], dtype=np.int32)
    nprng = np.random.default_rng(1)
    with L.Engine([n_hap]) as eng:
        eng.upload_allele_rows(0, panel, np.array(comp, dtype=np.uint8))
        eng.build_read_masks(flat, off)
        assert eng.n_reads == 120
        for n in (3, 4, 7):
            idx, doff = uniform_draws(nprng, 120, n, 12)
            for d in range(12):
                draw = idx[d * n:(d + 1) * n]
                assert eng.joint_counts(draw)[0].tolist()[1:] == O.joint_counts([masks[i] for i in draw])[1:]


def test_error_behaviour():
    rng = np.random.default_rng(2)
    rows = ld_rows(rng, 30, 100)
    with L.Engine([100]) as eng:
        with pytest.raises(L.LdpgtaError):                    # masks not uploaded
            eng.likelihoods_batch(L.HOMOGENEOUS, [0, 1], [0, 2])
        eng.upload_read_masks(0, rows)
        with pytest.raises(L.LdpgtaError, match='outside'):
            eng.likelihoods_batch(L.HOMOGENEOUS, [0, 30], [0, 2])
        with pytest.raises(L.LdpgtaError, match='supported range'):
            eng.likelihoods_batch(L.HOMOGENEOUS, list(range(19)), [0, 19])
        with pytest.raises(L.LdpgtaError, match='two'):
            eng.likelihoods_batch(L.RECENT_ADMIXTURE, [0, 1], [0, 2])
        assert eng.likelihoods_batch(L.HOMOGENEOUS, [], [0]).shape == (0, 4)
    with L.Engine([100, 100]) as eng:
        eng.upload_read_masks(0, rows)
        eng.upload_read_masks(1, rows)
        with pytest.raises(L.LdpgtaError, match='likelihoods5'):
            eng.likelihoods_batch(L.DISTANT_ADMIXTURE, list(range(5)), [0, 5], proportions=[0.5, 0.5])
    with L.Engine([70000]) as eng:                             # more than 65535 haplotypes: 32-bit count tables only
        big = ld_rows(rng, 20, 70000)
        eng.upload_read_masks(0, big)
        idx, off = uniform_draws(rng, 20, 6, 4)
        assert max_rel(eng.likelihoods_batch(L.HOMOGENEOUS, idx, off), C.batch([big], [70000], idx, off, 0)) < LIK_TOL
        idx, off = uniform_draws(rng, 20, 4, 4)
        assert max_rel(eng.likelihoods_batch(L.HOMOGENEOUS, idx, off), C.batch([big], [70000], idx, off, 0)) < LIK_TOL
        with pytest.raises(L.LdpgtaError, match='shared memory'):            # 15 rows of 70000 bits do not fit one SM
            eng.likelihoods_batch(L.HOMOGENEOUS, list(range(15)), [0, 15])


# ----------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full shapes
# ----------------------------------------------------------------------------------------

def test_full_size_properties_n4_genome_scale():
    """ 5008 haplotypes, ~29k windows x 32 draws of 4 reads (config 2 shape): sampled draws against
    the oracle, determinism, invariance to the order of reads inside a draw, and the bounds
    0 <= likelihood <= 1, monosomy = f(all reads). """
    rng = np.random.default_rng(2024)
    n_hap, n_reads, n_draws = 5008, 40000, 28700 * 32
    rows = ld_rows(rng, n_reads, n_hap)
    win = rng.integers(0, n_reads - 150, size=n_draws // 32)
    # four distinct reads out of the window's 150: a random start plus strictly increasing steps (< 150 in total)
    steps = np.concatenate([rng.integers(0, 30, size=(n_draws, 1)), rng.integers(1, 40, size=(n_draws, 3))], axis=1)
    idx = (np.repeat(win, 32)[:, None] + np.cumsum(steps, axis=1)).astype(np.int32)
    idx = np.take_along_axis(idx, np.argsort(rng.random((n_draws, 4)), axis=1), axis=1)    # draws are unordered
    off = np.arange(0, 4 * n_draws + 1, 4, dtype=np.int64)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        a = eng.likelihoods_batch(L.HOMOGENEOUS, idx.ravel(), off)
        b = eng.likelihoods_batch(L.HOMOGENEOUS, idx.ravel(), off)
        assert np.array_equal(a, b)                                            # deterministic
        perm = eng.likelihoods_batch(L.HOMOGENEOUS, idx[:, ::-1].copy().ravel(), off)
        assert max_rel(a, perm) < 1e-14                                       # symmetric in the reads
        assert np.all(a >= 0) and np.all(a <= 1)
        sample = rng.choice(n_draws, size=3000, replace=False)
        sidx = idx[sample].ravel()
        want = C.batch([rows], [n_hap], sidx, np.arange(0, 4 * 3000 + 1, 4, dtype=np.int64), 0)
        assert max_rel(a[sample], want) < LIK_TOL
        f_all = np.array([C.joint_counts(rows[idx[d]])[-1] for d in sample[:200]]) / n_hap
        assert np.array_equal(a[sample[:200], 0], f_all)                      # monosomy = joint frequency of all reads


def test_joint_count_table_properties_n16():
    """ 2^16 subsets: monotone under adding a read, singletons = row popcounts, duplicate read leaves
    the intersection unchanged. """
    rng = np.random.default_rng(5)
    n_hap = 5008
    rows = ld_rows(rng, 16, n_hap, founders=2, density=0.9)
    rows[7] = rows[3]                                                         # a duplicated read
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        F = eng.joint_counts(np.arange(16, dtype=np.int32))[0].astype(np.int64)
    pop = np.array([sum(int(w).bit_count() for w in r) for r in rows])
    assert np.array_equal(F[1 << np.arange(16)], pop)
    S = np.arange(1, 1 << 16)
    for i in range(16):
        with_i = S | (1 << i)
        assert np.all(F[with_i] <= F[S])
    has3 = S[(S >> 3) & 1 == 1]
    assert np.array_equal(F[has3 | (1 << 7)], F[has3])
    assert F[0] == 0


def test_window_stats_against_python_statistics():
    rng = np.random.default_rng(8)
    counts = rng.integers(2, 33, size=200)
    counts[0] = 32
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    lik = rng.random((off[-1], 4)) * 1e-3
    lik[rng.random(off[-1]) < 0.05, 3] = 0.0                                   # sentinels
    lik[rng.random(off[-1]) < 0.05, 2] = 0.0
    likd = {(k, k): tuple(O.likelihoods_tuple(*lik[i]) for i in range(off[k], off[k + 1])) for k in range(200)}
    want = O.statistics(likd, {w: tuple(range(10)) for w in likd})
    n_cols = int(counts.min())
    with L.Engine([1]) as eng:
        llr, mean_var, partials = eng.window_stats(lik, off, n_cols)
    chrom = L.finish_chromosome(partials, n_cols, int(counts[0]))
    for p, pair in enumerate(L.LLR_PAIRS):
        for k, w in enumerate(likd):
            m, v = want['LLRs_per_genomic_window'][pair][w]
            assert abs(mean_var[p, k, 0] - m) <= 1e-13 * max(1, abs(m))
            assert abs(mean_var[p, k, 1] - v) <= 1e-12 * max(1, abs(v))
        c = want['LLRs_per_chromosome'][pair]
        assert abs(chrom[p, 0] - c['mean_of_mean']) <= STAT_TOL * max(1, abs(c['mean_of_mean']))
        assert abs(chrom[p, 1] - c['std_of_mean']) <= STAT_TOL * max(1, abs(c['std_of_mean']))
        assert chrom[p, 2] == c['fraction_of_negative_LLRs']
        # the device forms log(y / x) without rounding the quotient first (common.cuh: log_ratio); math.log(y / x) carries
        # that rounding: up to 2^-53 absolute, which is all of the difference when y ~ x
        ref_llr = np.array([O.LLR(getattr(l, pair[0]), getattr(l, pair[1])) for ls in likd.values() for l in ls])
        assert np.all(np.abs(llr[p] - ref_llr) <= 1e-15 * np.abs(ref_llr) + 1.2e-16)
    # sharding: partials of two halves of the chromosome add up to the whole (the only cross-GPU exchange)
    with L.Engine([1]) as eng:
        h = 100
        _, _, p1 = eng.window_stats(lik[:off[h]], off[:h + 1], n_cols)
        _, _, p2 = eng.window_stats(lik[off[h]:], off[h:] - off[h], n_cols)
    both = L.finish_chromosome(p1 + p2, n_cols, int(counts[0]))
    assert np.max(np.abs(both - chrom)) < 1e-12


def test_llr_log_ratio_accuracy():
    """ LLR = log(y / x) (ANEUPLOIDY_TEST.py:78-90) on the device: one division for quotient and atanh argument
    (common.cuh: log_ratio).  Against a long-double value of the true logarithm it stays within 2.5 ulp -- relative accuracy
    also where y ~ x -- and it is never further from the reference's math.log(y / x) than that value's own error bound
    (1 ulp + the 2^-53 of the rounded quotient).  Subnormal likelihoods take the library route; zeros give the sentinels. """
    rng = np.random.default_rng(31)
    n = 32 * 4000
    y = np.exp(-40.0 * rng.random(n))
    x = y * (1.0 + (rng.random(n) - 0.5) * 2.0 ** (-rng.integers(0, 45, size=n)))
    far = slice(0, n // 4)                                                          # far apart: |log| up to ~600
    y[far], x[far] = np.exp(-600.0 * rng.random(n // 4)), np.exp(-600.0 * rng.random(n // 4))
    edge = slice(n // 4, n // 4 + 4000)                                             # at the range-reduction boundaries
    x[edge] = y[edge] * (np.where(rng.random(4000) < 0.5, 2.0 ** 0.5, 0.5 ** 0.5) + 1e-7 * (rng.random(4000) - 0.5))
    x[-64:-32] = y[-64:-32]                                                         # equal: exactly 0
    y[-32:-16] = 5e-324 * rng.integers(1, 1000, size=16)                           # subnormal numerators
    x[-16:] = 5e-324 * rng.integers(1, 1000, size=16)                              # subnormal denominators
    lik = np.ones((n, 4))
    lik[:, 3], lik[:, 2] = y, x                                                    # pair 0 = (BPH, SPH)
    off = np.arange(0, n + 1, 32, dtype=np.int64)
    with L.Engine([1]) as eng:
        llr, _, _ = eng.window_stats(lik, off, 32)
    got = llr[0]
    yl, xl = y.astype(np.longdouble), x.astype(np.longdouble)
    near = (y > 0.5 * x) & (y < 2.0 * x)
    truth = np.where(near, np.log1p((yl - xl) / xl), np.log(yl / xl))
    ulp = np.spacing(np.abs(truth.astype(np.float64)))
    err = np.abs(got.astype(np.longdouble) - truth).astype(np.float64)
    normal = (y >= 2.3e-308) & (x >= 2.3e-308)
    assert np.all(err[normal] <= 2.5 * ulp[normal]), float(np.max(err[normal] / ulp[normal]))      # 1.99 on a host model of the routine over 2e7 pairs
    # subnormal likelihoods: the library's division and logarithm, i.e. the reference's own arithmetic (a subnormal quotient
    # has lost bits, an overflowing one gives inf, exactly as CPython's y / x does)
    with np.errstate(over='ignore', divide='ignore'):
        ref = np.log(y / x)
    sub = ~normal
    assert np.array_equal(np.isinf(got[sub]), np.isinf(ref[sub]))
    fin = sub & np.isfinite(ref)
    assert np.all(np.abs(got[fin] - ref[fin]) <= 2.0 * np.spacing(np.abs(ref[fin])))
    assert np.all(np.abs(got[normal] - ref[normal]) <= 3.0 * ulp[normal] + 2.0 ** -53)
    assert np.all(got[-64:-32] == 0.0)
    assert np.all(llr[1] == 0.0)                                                   # (disomy, monosomy) = log(1 / 1)
    # sentinels (:83-85)
    lik2 = np.array([[0.0, 1.0, 0.0, 1.0], [1.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 0.0]])
    with L.Engine([1]) as eng:
        s, _, _ = eng.window_stats(lik2, np.array([0, 3], dtype=np.int64), 3)
    assert s[0].tolist() == [1.23456789, -1.23456789, 0.0] and s[1].tolist() == [1.23456789, -1.23456789, 0.0]


def test_same_seed_reproduces_reference_without_replay(golden_inputs):
    """ Drop-in property: with the reference's PYTHONHASHSEED and random seed, bootstrap() draws the same
    read subsets as the reference did (same random.sample call sequence) and returns its likelihoods. """
    import json, os, subprocess, sys
    from conftest import GOLDEN, ROOT
    code = r'''
import bz2, pickle, random, sys, json
sys.path.insert(0, %r)
import ldpgta_b200 as L
inp = pickle.load(bz2.open(%r + '/chrom_inputs.p.bz2', 'rb'))
out = pickle.load(bz2.open(%r + '/chrom_homog_m4.p.bz2', 'rb'))
cfg = out['cfg']
hap = {g: inp['hap_tab'][g] for g in cfg['groups']}
nh = {g: inp['number_of_haplotypes'][g] for g in cfg['groups']}
random.seed(2021)
lik, windows, _ = L.bootstrap(inp['obs_tabs'][cfg['scenario']], inp['leg_tab'], hap, nh, cfg['makeup'], L.ImplicitModels(),
                              cfg['window_size'], cfg['subsamples'], cfg['offset'], cfg['min_reads'], cfg['max_reads'],
                              cfg['min_score'], cfg['min_HF'])
same_windows = {w: tuple(r) for w, r in windows.items()} == {w: tuple(r) for w, r in out['windows'].items()}
worst = 0.0
for w, ls in out['likelihoods'].items():
    for mine, ref in zip(lik[w], ls):
        for a, b in zip(mine, ref):
            worst = max(worst, abs(a - b) / max(abs(a), abs(b), 1e-300))
print(json.dumps({'same_windows': same_windows, 'worst': worst, 'n': sum(len(v) for v in lik.values())}))
''' % (ROOT, GOLDEN, GOLDEN)
    env = dict(os.environ, PYTHONHASHSEED='0')
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    got = json.loads(res.stdout.strip().splitlines()[-1])
    assert got['same_windows'] and got['n'] == 392
    assert got['worst'] < LIK_TOL, got


@pytest.mark.parametrize('name', CHROM_CONFIGS)
def test_read_scores_on_device_match_reference(name, golden_inputs):
    """ build_score_dict (ANEUPLOIDY_TEST.py:151-189) on the GPU: exact integer scores, including the
    reference's last-alpha weighting for unequal group sizes. """
    args, out = chrom_case(name, golden_inputs)
    cfg = args['cfg']
    model = driver.make_examiner(args['obs_tab'], args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'],
                                 args['ancestral_makeup'], L.ImplicitModels())
    try:
        obs = driver.supporting_dictionaries(args['obs_tab'], args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'],
                                             args['ancestral_makeup'], cfg['min_HF'], examine=model)
        assert obs.score == out['score']
        assert list(obs.score) == list(out['score'])
    finally:
        model.close()


@pytest.mark.parametrize('name', ('homog_m4', 'distant_m4'))
def test_read_scores_of_long_reads(name, golden_inputs):
    """ A read that overlaps 13..15 common SNPs (long or paired reads): the reference enumerates all 2^k haplotype
    combinations (ANEUPLOIDY_TEST.py:178-184); the device spreads them over many warps.  Checked against the oracle's
    restatement of build_score_dict, which the golden chromosome vectors pin to the reference. """
    args, out = chrom_case(name, golden_inputs)
    cfg = args['cfg']
    obs = list(args['obs_tab'])
    seen, extra = set(), []
    for k, (first, step) in enumerate(((40, 1), (300, 1), (900, 2))):       # neighbouring SNPs: in LD, so combinations stay populated
        picked = []
        for pos, _rid, base in obs[first::step]:
            if pos not in seen and len(picked) < 13 + k:
                picked.append((pos, base)); seen.add(pos)
        extra += [(pos, 'long.read.%d' % k, base) for pos, base in picked]
    obs_long = sorted(obs + extra, key=lambda r: r[0])
    sd = O.SupportingDictionaries(obs_long, args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'], args['ancestral_makeup'], cfg['min_HF'])
    want = sd.score
    alpha = {g: sd.ancestral_makeup[g] / n for g, n in sd.number_of_haplotypes_per_group.items()}
    qualifying = [sum(0.01 <= sum(a * sd.hap_tab_per_group[g][pos].bit_count() for g, a in alpha.items()) <= 0.99 for pos, _ in sd.reads['long.read.%d' % k])
                  for k in range(3)]
    assert max(qualifying) >= 13, qualifying                   # beyond what one warp enumerates: the chunked kernel ran
    model = driver.make_examiner(obs_long, args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'], args['ancestral_makeup'], L.ImplicitModels())
    try:
        got = driver.supporting_dictionaries(obs_long, args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'],
                                             args['ancestral_makeup'], cfg['min_HF'], examine=model).score
    finally:
        model.close()
    assert got == want
    assert any(want['long.read.%d' % k] > 0 for k in range(3))


def test_repeatability_stress_multi_group():
    """ Races show up as run-to-run differences: many CTAs per SM, two groups, chunked tables; every batch is
    evaluated three times and must be bit-identical, and a slice is checked against the oracle. """
    rng = np.random.default_rng(77)
    sizes = (2504, 2504)
    rows = [ld_rows(rng, 300, s) for s in sizes]
    with L.Engine(list(sizes)) as eng:
        for g in range(2):
            eng.upload_read_masks(g, rows[g])
        for kind, prop, ck in ((L.RECENT_ADMIXTURE, None, 1), (L.DISTANT_ADMIXTURE, [0.3, 0.7], 2)):
            for n, nd in ((4, 5000), (6, 3000), (9, 1500), (12, 600), (14, 300)):
                idx, off = uniform_draws(rng, 300, n, nd)
                runs = [eng.likelihoods_batch(kind, idx, off, proportions=prop) for _ in range(3)]
                assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2]), (kind, n)
                k = 12 if n <= 9 else 3
                want = C.batch(_common_width(rows), list(sizes), idx[:k * n], off[:k + 1], ck, prop)
                assert max_rel(runs[0][:k], want) < LIK_TOL, (kind, n)


def test_aneuploidy_test_files_roundtrip(tmp_path, golden_inputs):
    """ The driver's file boundary (ANEUPLOIDY_TEST.py:368-429): OBS/LEG/HAP pickles in (plain, gz, bz2), the
    two-pickle *.LLR.p out, `info` keys, and likelihoods_tuple records that unpickle under the reference's
    module names. """
    import bz2, gzip, pickle, subprocess, sys
    from conftest import ROOT
    args, out = chrom_case('homog_m4', golden_inputs)
    cfg = args['cfg']
    d = tmp_path
    with open(d / 'sample.chr21.obs.p', 'wb') as f:
        pickle.dump(args['obs_tab'], f); pickle.dump({'chr_id': 'chr21', 'depth': 0.45}, f)
    with gzip.open(d / 'panel.legend.gz', 'wb') as f:
        pickle.dump(args['leg_tab'], f)
    with bz2.open(d / 'panel.hap.bz2', 'wb') as f:
        pickle.dump(args['hap_tab'], f); pickle.dump(args['number_of_haplotypes'], f)
    code = r'''
import sys, pickle, json
sys.path.insert(0, %r)                       # flat drop-in: the directory with the reference's module names
from ANEUPLOIDY_TEST import aneuploidy_test
lik, info = aneuploidy_test(%r, %r, %r, {'G0'}, %d, %d, %d, %d, %d, %d, %r, '', 'gz', seed=2021, output_dir=%r)
print(json.dumps({'windows': len(lik), 'keys': sorted(info), 'stat_keys': sorted(info['statistics'])}))
''' % (str(ROOT) + '/ldpgta_b200', str(d / 'sample.chr21.obs.p'), str(d / 'panel.legend.gz'), str(d / 'panel.hap.bz2'),
       cfg['window_size'], cfg['subsamples'], cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'],
       cfg['min_HF'], str(d / 'results'))
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300, cwd=str(d),
                         env=dict(__import__('os').environ, PYTHONHASHSEED='0'))
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    got = json.loads(res.stdout.strip().splitlines()[-1])
    assert got['windows'] == len(out['likelihoods'])
    for k in ('ancestral_makeup', 'window_size', 'subsamples', 'offset', 'min_reads', 'max_reads', 'min_score', 'min_HF',
              'statistics', 'chr_id', 'depth'):
        assert k in got['keys']
    for k in ('matched_alleles', 'runtime', 'num_of_windows', 'LLRs_per_chromosome', 'LLRs_per_genomic_window',
              'reads_mean', 'window_size_mean'):
        assert k in got['stat_keys']
    path = d / 'results' / 'sample.chr21.LLR.p.gz'
    assert path.exists()
    raw = gzip.open(path, 'rb').read()
    assert b'HOMOGENOUES_MODELS' in raw and b'ldpgta' not in raw          # records carry the reference's module name
    sys.path.insert(0, str(ROOT) + '/ldpgta_b200')
    with gzip.open(path, 'rb') as f:
        lik = pickle.load(f); info = pickle.load(f)
    assert list(lik) == list(out['likelihoods'])
    # same seed and hash seed as the golden run: the draws and therefore the likelihoods are the reference's
    worst = max(max_rel(a, b) for w in lik for a, b in zip(lik[w], out['likelihoods'][w]))
    assert worst < LIK_TOL
    # statistics() comes from the same device pass as the likelihoods: every pair, per chromosome and per window, against
    # the reference's recorded statistics
    for pair in L.LLR_PAIRS:
        ref, mine = out['statistics']['LLRs_per_chromosome'][pair], info['statistics']['LLRs_per_chromosome'][pair]
        for key in ('mean_of_mean', 'std_of_mean', 'fraction_of_negative_LLRs'):
            assert abs(mine[key] - ref[key]) <= STAT_TOL * max(1e-3, abs(ref[key])), (pair, key)
        for w, (m, v) in out['statistics']['LLRs_per_genomic_window'][pair].items():
            gm, gv = info['statistics']['LLRs_per_genomic_window'][pair][w]
            assert abs(gm - m) <= STAT_TOL * max(1e-3, abs(m)) and abs(gv - v) <= STAT_TOL * max(1e-3, abs(v))
    for key in ('num_of_windows', 'reads_mean', 'reads_std', 'window_size_mean', 'window_size_std'):
        assert info['statistics'][key] == out['statistics'][key]


def test_resident_panel_gather_equals_host_packing(golden_testmodule, golden_inputs):
    """ Next row f3: the panel stays on the device and build_hap_dict becomes a gather
    (ldpgta_panel_* + ldpgta_gather_allele_rows); results are identical to packing on the host. """
    g = golden_testmodule
    obs, leg, hap, nh = g['obs_tab'], g['leg_tab'], g['hap_tab'], g['number_of_haplotypes']
    panel1 = L.ResidentPanel(leg, {'group0': hap['group0']}, nh)
    panel2 = L.ResidentPanel(leg, hap, nh)
    a = L.homogeneous(obs, leg, {'group0': hap['group0']}, {'group0': nh['group0']}, L.ImplicitModels())
    b = L.homogeneous(obs, leg, {'group0': hap['group0']}, {'group0': nh['group0']}, L.ImplicitModels(), panel=panel1)
    c = L.recent_admixture(obs, leg, hap, nh, L.ImplicitModels())
    d = L.recent_admixture(obs, leg, hap, nh, L.ImplicitModels(), panel=panel2)
    for rec in g['records']:
        al = rec['alleles']
        assert a.joint_frequencies_combo(al, normalize=False) == b.joint_frequencies_combo(al, normalize=False)
        assert tuple(a.get_likelihoods(al)) == tuple(b.get_likelihoods(al))
        assert tuple(c.get_likelihoods(al)) == tuple(d.get_likelihoods(al))
    for m in (a, b, c, d):
        m.close()
    # several samples share one resident panel through the driver
    args, out = chrom_case('recent_m4', golden_inputs)
    cfg = args['cfg']
    shared = L.ResidentPanel(args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'])
    for _ in range(2):
        lik, _, matches = L.bootstrap(args['obs_tab'], args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'],
                                      args['ancestral_makeup'], L.ImplicitModels(), cfg['window_size'], cfg['subsamples'],
                                      cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'], cfg['min_HF'],
                                      replay=out['draws'], panel=shared)
        assert matches == out['fraction_of_matches']
        worst = max(max_rel(x, y) for w in lik for x, y in zip(lik[w], out['likelihoods'][w]))
        assert worst < LIK_TOL
    for p in (panel1, panel2, shared):
        p.close()
    # distant admixture: the shared panel holds its groups in the order of the ancestral makeup (G1 first here)
    args, out = chrom_case('distant_m6', golden_inputs)
    cfg = args['cfg']
    order = list(args['ancestral_makeup'])
    shared = L.ResidentPanel(args['leg_tab'], {g: args['hap_tab'][g] for g in order}, args['number_of_haplotypes'])
    lik, _, matches = L.bootstrap(args['obs_tab'], args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'],
                                  args['ancestral_makeup'], L.ImplicitModels(), cfg['window_size'], cfg['subsamples'],
                                  cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'], cfg['min_HF'],
                                  replay=out['draws'], panel=shared)
    assert matches == out['fraction_of_matches']
    assert max(max_rel(x, y) for w in lik for x, y in zip(lik[w], out['likelihoods'][w])) < LIK_TOL
    shared.close()


def test_command_line_interface(tmp_path, golden_inputs):
    """ `python ANEUPLOIDY_TEST.py OBS LEG HAP G0 -s 8 ...` (the reference's CLI, ANEUPLOIDY_TEST.py:440-493). """
    import gzip, os, pickle, subprocess, sys
    from conftest import ROOT
    args, out = chrom_case('homog_m4', golden_inputs)
    d = tmp_path
    with open(d / 'cli.chr21.obs.p', 'wb') as f:
        pickle.dump(args['obs_tab'], f); pickle.dump({'chr_id': 'chr21', 'depth': 0.45}, f)
    with open(d / 'panel.legend.p', 'wb') as f:
        pickle.dump(args['leg_tab'], f)
    with open(d / 'panel.hap.p', 'wb') as f:
        # the reference's CLI only accepts alphabetic group names (ANEUPLOIDY_TEST.py:480-488): G0 -> EUR
        pickle.dump({'EUR': args['hap_tab']['G0']}, f); pickle.dump({'EUR': args['number_of_haplotypes']['G0']}, f)
    script = os.path.join(str(ROOT), 'ldpgta_b200', 'ANEUPLOIDY_TEST.py')
    res = subprocess.run([sys.executable, script, str(d / 'cli.chr21.obs.p'), str(d / 'panel.legend.p'), str(d / 'panel.hap.p'),
                          'EUR', '-s', '8', '-m', '6', '-M', '4', '-C', 'gz'], capture_output=True, text=True, timeout=300, cwd=str(d))
    assert res.returncode == 0, res.stderr[-2000:]
    assert 'Chromosome-wide LLR between BPH and SPH' in res.stdout
    with gzip.open(d / 'results' / 'cli.chr21.LLR.p.gz', 'rb') as f:
        sys.path.insert(0, os.path.join(str(ROOT), 'ldpgta_b200'))
        lik = pickle.load(f); info = pickle.load(f)
    assert list(lik) == list(out['likelihoods']) and info['subsamples'] == 8 and info['ancestral_makeup'] == {'EUR'}
    bad = subprocess.run([sys.executable, script, 'a', 'b', 'c', 'G0'], capture_output=True, text=True, timeout=300)
    assert bad.returncode != 0 and 'Invalid argument supplied for ancestral makeup' in bad.stderr


def test_uniform_host_api_pipeline():
    """ ldpgta_likelihoods_uniform: draws of one size as a 2-D array, page-locked and pageable buffers, equal to the
    ragged API and to the oracle; a bad index is caught by the device-side validation and named in the error. """
    rng = np.random.default_rng(77)
    n_hap, n_reads = 5008, 3000
    rows = ld_rows(rng, n_reads, n_hap)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        for n, n_draws in ((4, 200000), (3, 70001), (6, 66000), (9, 1500)):
            idx = rng.integers(0, n_reads, size=(n_draws, n), dtype=np.int32)
            idx[:, 1:] = (idx[:, :1] + np.arange(1, n, dtype=np.int32) * 7) % n_reads       # distinct reads in a draw
            off = np.arange(0, n_draws * n + 1, n, dtype=np.int64)
            ragged = eng.likelihoods_batch(L.HOMOGENEOUS, idx.reshape(-1), off)
            pageable = eng.likelihoods_uniform(L.HOMOGENEOUS, idx)
            pin_i, pin_o = L.PinnedArray((n_draws, n), np.int32), L.PinnedArray((n_draws, 4), np.float64)
            pin_i.array[:] = idx
            pinned = eng.likelihoods_uniform(L.HOMOGENEOUS, pin_i.array, out=pin_o.array)
            via_offsets = eng.likelihoods_batch(L.HOMOGENEOUS, pin_i.array.reshape(-1), off, out=np.zeros_like(pin_o.array))
            assert np.array_equal(ragged, pageable) and np.array_equal(ragged, pinned) and np.array_equal(ragged, via_offsets)
            sample = rng.choice(n_draws, 300, replace=False)
            ref = C.batch([rows], [n_hap], idx[sample].reshape(-1), np.arange(0, 300 * n + 1, n), 0)
            assert max_rel(ragged[sample], ref) < LIK_TOL
            # one bad index somewhere in the middle: reported with its position, from both kinds of buffer
            k = n_draws // 2 * n + 1
            for bad in (n_reads, -1):
                pin_i.array.reshape(-1)[k] = bad
                with pytest.raises(L.LdpgtaError, match=r'read index %d at %d is outside' % (bad, k)):
                    eng.likelihoods_uniform(L.HOMOGENEOUS, pin_i.array, out=pin_o.array)
                with pytest.raises(L.LdpgtaError, match=r'read index %d at %d is outside' % (bad, k)):
                    eng.likelihoods_uniform(L.HOMOGENEOUS, np.array(pin_i.array))
                if n_draws >= 65536:
                    with pytest.raises(L.LdpgtaError, match=r'read index %d at %d is outside' % (bad, k)):
                        eng.likelihoods_batch(L.HOMOGENEOUS, pin_i.array.reshape(-1), off, out=pin_o.array)
            pin_i.array[:] = idx
            assert np.array_equal(eng.likelihoods_uniform(L.HOMOGENEOUS, pin_i.array, out=pin_o.array), ragged)
            # offsets that look uniform at both ends but are not: the pipelined path hands over to the regrouping path
            if n_draws >= 65536 and n == 4:
                off2 = off.copy()
                off2[n_draws // 3] += 1                       # draw sizes 5 and 3 next to each other
                got = eng.likelihoods_batch(L.HOMOGENEOUS, pin_i.array.reshape(-1), off2, out=pin_o.array)
                ref2 = eng.likelihoods_batch(L.HOMOGENEOUS, idx.reshape(-1), off2)
                assert np.array_equal(got, ref2)
        with pytest.raises(L.LdpgtaError, match='supported range'):
            eng.likelihoods_uniform(L.HOMOGENEOUS, np.zeros((4, 1), np.int32))
        with pytest.raises(ValueError):
            eng.likelihoods_uniform(L.HOMOGENEOUS, np.zeros(8, np.int32))
        assert eng.likelihoods_uniform(L.HOMOGENEOUS, np.zeros((0, 4), np.int32)).shape == (0, 4)


def test_host_pipeline_graph_replay():
    """ Calls that repeat the previous call's shape and page-locked buffers replay one CUDA graph: refilled buffers must
    give the new results, other shapes in between must not disturb it, bad indices must still be caught, and the table
    kernel's global scratch (17 reads) must survive a reallocation between replays. """
    rng = np.random.default_rng(99)
    n_hap, n_reads = 2504, 400
    rows = ld_rows(rng, n_reads, n_hap)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)

        def draws(n_draws, n):
            idx = rng.integers(0, n_reads, size=(n_draws, 1), dtype=np.int32) + np.arange(n, dtype=np.int32) * 3
            return (idx % n_reads).astype(np.int32)

        def check(pin_i, pin_o, idx, n_check=64):
            pin_i.array[:] = idx
            got = eng.likelihoods_uniform(L.HOMOGENEOUS, pin_i.array, out=pin_o.array)
            pick = rng.choice(len(idx), min(n_check, len(idx)), replace=False)
            n = idx.shape[1]
            ref = C.batch([rows], [n_hap], idx[pick].reshape(-1), np.arange(0, len(pick) * n + 1, n), 0)
            assert max_rel(got[pick], ref) < LIK_TOL

        a_i, a_o = L.PinnedArray((50000, 4), np.int32), L.PinnedArray((50000, 4), np.float64)
        b_i, b_o = L.PinnedArray((3000, 9), np.int32), L.PinnedArray((3000, 4), np.float64)
        c_i, c_o = L.PinnedArray((40, 17), np.int32), L.PinnedArray((40, 4), np.float64)
        d_i, d_o = L.PinnedArray((400, 17), np.int32), L.PinnedArray((400, 4), np.float64)
        for _ in range(4):                                   # eager, eager/capture, replay, replay -- new contents every time
            check(a_i, a_o, draws(50000, 4))
        launches = eng.launch_count
        check(a_i, a_o, draws(50000, 4))
        assert eng.launch_count > launches                   # replays are counted as kernel launches too
        assert eng.graph_replays >= 3
        for _ in range(3):
            check(b_i, b_o, draws(3000, 9))
        check(a_i, a_o, draws(50000, 4))                     # back to the first shape: its graph was replaced, still correct
        a_i.array[:] = draws(50000, 4)
        a_i.array[31234, 2] = n_reads + 5
        for _ in range(3):                                   # the failing call is eager, captured or replayed: always caught
            with pytest.raises(L.LdpgtaError, match='read index %d at %d is outside' % (n_reads + 5, 31234 * 4 + 2)):
                eng.likelihoods_uniform(L.HOMOGENEOUS, a_i.array, out=a_o.array)
        check(a_i, a_o, draws(50000, 4))
        for _ in range(3):
            check(c_i, c_o, draws(40, 17), n_check=6)        # tables in global scratch, one slice per resident CTA (40 CTAs)
        check(d_i, d_o, draws(400, 17), n_check=6)           # more CTAs: the scratch is reallocated
        for _ in range(3):
            check(c_i, c_o, draws(40, 17), n_check=6)        # a stale graph would write through the freed scratch
        assert eng.graph_replays >= 8
