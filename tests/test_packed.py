"""
The array-native route (ldpgta_b200/packed.py) against the golden vectors made by running the
reference: matching of observations to the legend, reads as CSR, windows (fixed, offset, adaptive) on the CPU;
read scores, replayed draws and the whole run on the GPU.
"""
import numpy as np
import pytest

import ldpgta_oracle as O
from conftest import CHROM_CONFIGS, chrom_case, rel_err


def packed_inputs(name, golden_inputs):
    from ldpgta_b200.packed import PackedLegend, pack_observations
    args, out = chrom_case(name, golden_inputs)
    legend = PackedLegend(args['leg_tab'])
    return args, out, legend, pack_observations(args['obs_tab'], legend)


@pytest.mark.parametrize('name', CHROM_CONFIGS)
def test_pack_observations_and_windows_cpu(name, golden_inputs):
    from ldpgta_b200.packed import build_windows
    args, out, legend, po = packed_inputs(name, golden_inputs)
    cfg = args['cfg']
    g0 = cfg['groups'][0]
    assert po.fraction_of_matches == out['fraction_of_matches'][g0]
    # reads: same read ids, same alleles in the same order as build_reads_dict
    leg_dict = {l[1]: (l[2], l[3]) for l in args['leg_tab']}
    reads = {}
    for pos, rid, base in args['obs_tab']:
        if base in leg_dict.get(pos, ()):
            reads.setdefault(rid, []).append((pos, base))
    assert sorted(po.read_names) == sorted(reads) and set(out['score']) == set(reads)
    leg_pos = [l[1] for l in args['leg_tab']]
    leg = args['leg_tab']
    for r, rid in enumerate(po.read_names):
        got = []
        for a in po.read_alleles[po.read_offsets[r]:po.read_offsets[r + 1]]:
            l = leg[int(po.allele_rows[a])]
            got.append((l[1], l[2] if po.allele_complement[a] else l[3]))
        assert got == reads[rid]
    # windows from the golden scores: same borders in the same order, same read sets
    scores = np.array([out['score'][rid] for rid in po.read_names])
    win = build_windows(po.obs_pos, po.obs_read, scores, cfg['window_size'], cfg['offset'], cfg['min_reads'], cfg['min_score'])
    assert [tuple(b) for b in win.borders.tolist()] == list(out['windows'])
    for (lo, hi), ids in zip(zip(win.offsets[:-1], win.offsets[1:]), out['windows'].values()):
        assert {po.read_names[r] for r in win.reads[lo:hi]} == set(ids)


def test_packed_observations_roundtrip(tmp_path, golden_inputs):
    from ldpgta_b200.packed import PackedObservations
    args, out, legend, po = packed_inputs('recent_m4', golden_inputs)
    po.save(str(tmp_path / 'sample.npz'))
    back = PackedObservations.load(str(tmp_path / 'sample.npz'))
    assert back.read_names == po.read_names and back.fraction_of_matches == po.fraction_of_matches and back.n_snps == len(args['leg_tab'])
    for k in PackedObservations._arrays:
        assert np.array_equal(getattr(back, k), getattr(po, k)) and getattr(back, k).dtype == getattr(po, k).dtype


def test_legend_lookup_semantics():
    from ldpgta_b200.packed import PackedLegend, pack_observations
    leg = [('c', 10, 'A', 'G'), ('c', 20, 'C', 'T'), ('c', 20, 'G', 'T'), ('c', 5, 'T', 'TA')]       # duplicate position, unsorted
    legend = PackedLegend(leg)
    obs = [(5, 'r0', 'TA'), (10, 'r0', 'A'), (10, 'r1', 'G'), (20, 'r1', 'C'), (20, 'r2', 'G'), (20, 'r2', 'T'), (30, 'r3', 'A'), (10, 'r3', 'N')]
    with pytest.raises(ValueError):
        pack_observations(obs + [(5, 'r4', 'T')], legend)   # matched observations must be sorted by position
    obs = sorted(obs, key=lambda r: r[0])
    po = pack_observations(obs, legend)
    # the later legend row of position 20 wins: 'C' is a mismatch, 'G' is REF of row 2; unknown position and base are mismatches
    assert po.fraction_of_matches == 1 - 3 / 8
    alleles = {(int(r), int(c)) for r, c in zip(po.allele_rows, po.allele_complement)}
    assert alleles == {(3, 0), (0, 1), (0, 0), (2, 1), (2, 0)}
    assert po.read_names == ['r0', 'r1', 'r2'] and po.read_offsets.tolist() == [0, 2, 3, 5]


def test_sampling_without_replacement_is_uniform():
    from ldpgta_b200.packed import sample_without_replacement
    rng = np.random.default_rng(5)
    R = np.full(60000, 7)
    s = sample_without_replacement(rng, R, 3)
    assert s.min() == 0 and s.max() == 6 and all(len(set(row)) == 3 for row in s[:2000].tolist())
    counts = np.bincount(np.sort(s, axis=1) @ np.array([49, 7, 1]), minlength=343)
    seen = counts[counts > 0]
    assert len(seen) == 35 and abs(seen / len(R) - 1 / 35).max() < 0.004        # all 35 subsets, equally likely
    R = rng.integers(5, 400, size=5000)
    s = sample_without_replacement(rng, R, 4)
    assert (s < R[:, None]).all() and (s >= 0).all() and all(len(set(row)) == 4 for row in s.tolist())


@pytest.mark.gpu
@pytest.mark.parametrize('name', CHROM_CONFIGS)
def test_packed_sample_matches_reference(name, golden_inputs):
    import ldpgta_b200 as L
    from ldpgta_b200.packed import PackedSample, pack_observations
    args, out = chrom_case(name, golden_inputs)
    cfg = args['cfg']
    groups = list(cfg['makeup']) if isinstance(cfg['makeup'], dict) else list(cfg['groups'])
    hap = {g: args['hap_tab'][g] for g in groups}
    nhap = {g: args['number_of_haplotypes'][g] for g in args['number_of_haplotypes']}
    panel = L.ResidentPanel(args['leg_tab'], hap, nhap)
    with PackedSample(args['obs_tab'], panel, cfg['makeup'], cfg['min_HF']) as ps:
        assert ps.fraction_of_matches == out['fraction_of_matches']
        assert dict(zip(ps.read_names, ps.scores.tolist())) == out['score']
        with PackedSample(pack_observations(args['obs_tab'], panel.legend), panel, cfg['makeup'], cfg['min_HF']) as again:
            assert np.array_equal(again.scores, ps.scores) and again.read_names == ps.read_names      # from arrays: same sample
        win = ps.windows(cfg['window_size'], cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'])
        assert [tuple(b) for b in win.borders.tolist()] == list(out['windows'])
        # the reference's own draws, replayed as read rows
        row_of = {rid: r for r, rid in enumerate(ps.read_names)}
        by_k, ref = {}, {}
        for w, draws in out['draws'].items():
            for d, lik in zip(draws, out['likelihoods'][w]):
                by_k.setdefault(len(d), []).append([row_of[r] for r in d])
                ref.setdefault(len(d), []).append(lik)
        worst = 0.0
        for k, idx in by_k.items():
            got = ps.evaluate([(np.arange(len(idx)), np.array(idx, dtype=np.int32))], len(idx))
            for g, r in zip(got.tolist(), ref[k]):
                worst = max(worst, max(rel_err(a, b) for a, b in zip(g, r)))
        assert worst <= 1e-12, worst
        # the same draws through the device-resident pipeline (set_windows + replay_stats): the reference's likelihoods
        # to 1e-12 and the reference's statistics() -- per window and per chromosome -- to 1e-9
        flat = [row_of[r] for w in out['likelihoods'] for d in out['draws'][w] for r in d]
        res = ps.run(cfg['window_size'], cfg['subsamples'], cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'], replay=flat)
        assert [tuple(b) for b in res.borders.tolist()] == list(out['likelihoods'])
        ref_lik = np.array([l for w in out['likelihoods'] for l in out['likelihoods'][w]], dtype=np.float64)
        assert np.max(np.abs(res.likelihoods - ref_lik) / np.maximum(np.abs(ref_lik), 1e-300)) <= 1e-12
        gold = out['statistics']
        for p_i, pair in enumerate(L.LLR_PAIRS):
            for k, w in enumerate(out['likelihoods']):
                m, v = gold['LLRs_per_genomic_window'][pair][w]
                gm, gv = res.window_mean_var[p_i, k]
                assert abs(gm - m) <= 1e-9 * max(1e-3, abs(m)) and abs(gv - v) <= 1e-9 * max(1e-3, abs(v))
            for key, val in gold['LLRs_per_chromosome'][pair].items():
                assert abs(res.llr_per_chromosome()[pair][key] - val) <= 1e-9 * max(1e-3, abs(val)), (pair, key)
        stats_only = ps.run(cfg['window_size'], cfg['subsamples'], cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'],
                            replay=flat, want_likelihoods=False, want_draws=False)
        assert stats_only.likelihoods is None and np.array_equal(stats_only.chromosome, res.chromosome)
        # a full packed run: every window the reference evaluates is evaluated, draws are valid, statistics finite
        for mode in ({'rng': np.random.default_rng(1)}, {'seed': 5}):           # host-sampled and device-sampled draws
            res = ps.run(cfg['window_size'], cfg['subsamples'], cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'], **mode)
            lik, gw, frac = res.as_reference()
            assert list(lik) == list(out['likelihoods']) and list(gw) == list(out['windows'])
            assert [len(v) for v in lik.values()] == [len(v) for v in out['likelihoods'].values()]
            sizes = dict(zip(map(tuple, win.borders.tolist()), win.sizes.tolist()))
            for rows, idx in res.draws:
                assert all(len(set(d)) == len(d) for d in idx.tolist())
            for w, ids in gw.items():
                assert set(ids) == set(out['windows'][w])
            assert np.isfinite(res.likelihoods).all() and (res.likelihoods >= 0).all() and (res.likelihoods <= 1).all()
            stats = res.llr_per_chromosome()
            oracle_stats = O.statistics({w: [O.likelihoods_tuple(*l) for l in v] for w, v in lik.items()}, gw)
            for pair, s in stats.items():
                for key in ('mean_of_mean', 'std_of_mean', 'fraction_of_negative_LLRs'):
                    assert rel_err(s[key], oracle_stats['LLRs_per_chromosome'][pair][key]) <= 1e-9 or abs(s[key] - oracle_stats['LLRs_per_chromosome'][pair][key]) < 1e-15
    panel.close()


@pytest.mark.gpu
@pytest.mark.parametrize('name', ('homog_m4', 'recent_m6', 'distant_m6'))
def test_aneuploidy_test_packed_option(name, tmp_path, golden_inputs):
    """ aneuploidy_test(..., packed=True): same files, same `info` / statistics structure, the windows of the reference,
    statistics equal to the oracle's statistics() of the returned likelihoods. """
    import pickle
    from ldpgta_b200 import driver
    args, out = chrom_case(name, golden_inputs)
    cfg = args['cfg']
    d = tmp_path
    with open(d / 's.chr21.obs.p', 'wb') as f:
        pickle.dump(args['obs_tab'], f); pickle.dump({'chr_id': 'chr21', 'depth': 0.3}, f)
    with open(d / 'leg.p', 'wb') as f:
        pickle.dump(args['leg_tab'], f)
    with open(d / 'hap.p', 'wb') as f:
        pickle.dump(args['hap_tab'], f); pickle.dump(args['number_of_haplotypes'], f)
    lik, info = driver.aneuploidy_test(str(d / 's.chr21.obs.p'), str(d / 'leg.p'), str(d / 'hap.p'), cfg['makeup'], cfg['window_size'],
                                       cfg['subsamples'], cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'],
                                       cfg['min_HF'], '', 'unc', seed=3, output_dir=str(d / 'results'), packed=True)
    assert list(lik) == list(out['likelihoods'])
    S = info['statistics']
    assert S['matched_alleles'] == out['fraction_of_matches'] and S['num_of_windows'] == len(lik)
    gw = {w: ids for w, ids in out['windows'].items()}
    ref = O.statistics({w: [O.likelihoods_tuple(*l) for l in v] for w, v in lik.items()}, gw)
    for key in ('reads_mean', 'reads_std', 'window_size_mean', 'window_size_std'):
        assert rel_err(S[key], ref[key]) < 1e-12
    for pair, per_w in ref['LLRs_per_genomic_window'].items():
        for w, (m, v) in per_w.items():
            gm, gv = S['LLRs_per_genomic_window'][pair][w]
            assert abs(gm - m) <= 1e-12 * max(1, abs(m)) and abs(gv - v) <= 1e-9 * max(1e-3, abs(v))
        for key, val in ref['LLRs_per_chromosome'][pair].items():
            assert abs(S['LLRs_per_chromosome'][pair][key] - val) <= 1e-9 * max(1e-3, abs(val))
    with open(d / 'results' / 's.chr21.LLR.p', 'rb') as f:
        saved = pickle.load(f)
    assert list(saved) == list(lik) and type(next(iter(saved.values()))[0]).__name__ == 'likelihoods_tuple'


def philox4x32_10(c, k):
    """ numpy restatement of Philox4x32-10 (Salmon et al., SC'11): c uint32[n,4], k uint32[2] -> uint32[n,4] """
    c = c.astype(np.uint64).copy()
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32)


def sample_draws_reference(window_offsets, window_reads, draw_offsets, n, seed):
    """ what k_sample_draws must produce: Floyd's algorithm on Philox integers, draw by draw """
    n_draws = int(draw_offsets[-1])
    win = np.searchsorted(draw_offsets, np.arange(n_draws), side='right') - 1
    base, R = window_offsets[win], (window_offsets[win + 1] - window_offsets[win])
    d = np.arange(n_draws, dtype=np.uint64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    picks = np.empty((n_draws, n), dtype=np.int64)
    for i in range(n):
        if i % 2 == 0:
            ctr = np.stack([d & np.uint64(0xFFFFFFFF), d >> np.uint64(32), np.full(n_draws, i // 2, np.uint64), np.zeros(n_draws, np.uint64)], axis=1)
            rnd = philox4x32_10(ctr.astype(np.uint32), key).astype(np.uint64)
        x = (rnd[:, 3] << np.uint64(32) | rnd[:, 2]) if i % 2 else (rnd[:, 1] << np.uint64(32) | rnd[:, 0])
        top = R - n + i
        t = np.array([(int(a) * (int(b) + 1)) >> 64 for a, b in zip(x.tolist(), top.tolist())], dtype=np.int64)
        dup = (picks[:, :i] == t[:, None]).any(axis=1) if i else np.zeros(n_draws, bool)
        picks[:, i] = np.where(dup, top, t)
    return window_reads[base[:, None] + picks].astype(np.int32)


def test_philox_known_answers():
    """ Random123's published known-answer vectors for philox4x32-10 pin the restatement used as the sampler's oracle """
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == want


@pytest.mark.gpu
def test_device_draws_bit_exact_and_uniform():
    import ldpgta_b200 as L
    rng = np.random.default_rng(3)
    n_hap, n_reads = 331, 900
    rows = rng.integers(0, 2 ** 63, size=(n_reads, 6), dtype=np.uint64) | rng.integers(0, 2 ** 63, size=(n_reads, 6), dtype=np.uint64)
    rows[:, -1] &= np.uint64((1 << (n_hap - 320)) - 1)
    sizes = rng.integers(5, 60, size=40)
    sizes[:3] = (4, 5, 7)
    window_offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    window_reads = rng.permutation(n_reads)[:window_offsets[-1]].astype(np.int32) if window_offsets[-1] <= n_reads else rng.integers(0, n_reads, window_offsets[-1]).astype(np.int32)
    with L.Engine([n_hap]) as eng:
        eng.upload_read_masks(0, rows)
        for n, seed in ((4, 12345), (3, 2 ** 63 + 11), (7, 1)):
            per_window = np.where(sizes >= n, rng.integers(0, 50, size=len(sizes)), 0)
            draw_offsets = np.concatenate([[0], np.cumsum(per_window)]).astype(np.int64)
            lik, idx = eng.bootstrap_uniform(L.HOMOGENEOUS, n, window_offsets, window_reads, draw_offsets, seed, want_draws=True)
            want = sample_draws_reference(window_offsets, window_reads, draw_offsets, n, seed)
            assert np.array_equal(idx, want)                                   # the sampler, bit for bit
            assert np.array_equal(lik, eng.likelihoods_uniform(L.HOMOGENEOUS, idx))     # and the likelihoods of exactly those draws
            lik2 = eng.bootstrap_uniform(L.HOMOGENEOUS, n, window_offsets, window_reads, draw_offsets, seed)
            assert np.array_equal(lik, lik2)
            assert not np.array_equal(idx, eng.bootstrap_uniform(L.HOMOGENEOUS, n, window_offsets, window_reads, draw_offsets, seed + 1, want_draws=True)[1])
        # uniformity over k-subsets: one window of 7 reads, 3 of them per draw, 35 subsets
        woff, wr = np.array([0, 7], dtype=np.int64), np.arange(100, 107, dtype=np.int32)
        doff = np.array([0, 140000], dtype=np.int64)
        _, idx = eng.bootstrap_uniform(L.HOMOGENEOUS, 3, woff, wr, doff, 77, want_draws=True)
        assert idx.min() == 100 and idx.max() == 106 and (np.diff(np.sort(idx, axis=1), axis=1) > 0).all()
        counts = np.bincount((np.sort(idx - 100, axis=1) @ np.array([49, 7, 1])), minlength=343)
        seen = counts[counts > 0]
        assert len(seen) == 35 and abs(seen / 140000 - 1 / 35).max() < 0.003
        # errors
        from ldpgta_b200._native import LdpgtaError
        with pytest.raises(LdpgtaError, match='fewer than'):
            eng.bootstrap_uniform(L.HOMOGENEOUS, 8, woff, wr, doff, 1)
        with pytest.raises(LdpgtaError, match='outside'):
            eng.bootstrap_uniform(L.HOMOGENEOUS, 3, woff, wr + 1000, doff, 1)


def test_plan_counts_matches_effective_number_of_subsamples():
    """ draws per window and reads per draw for all windows at once = effective_number_of_subsamples / pick_reads
    (ANEUPLOIDY_TEST.py:226-248), incl. windows with fewer than max_reads + 1 reads and comb(R, max_reads) < subsamples """
    from ldpgta_b200.packed import PackedSample, PackedWindows
    rng = np.random.default_rng(0)
    for min_reads, max_reads, subsamples in ((6, 4, 32), (3, 3, 5), (12, 8, 6), (4, 2, 1000), (18, 16, 32), (6, 4, 1), (2, 2, 40)):
        R = np.concatenate([np.arange(0, 40), rng.integers(0, 400, size=200)])
        offsets = np.concatenate([[0], np.cumsum(R)]).astype(np.int64)
        win = PackedWindows(np.zeros((len(R), 2), np.int64), offsets, np.zeros(int(offsets[-1]), np.int64))
        evaluated, draw_offsets, effN, k_of = PackedSample.plan_counts(win, min_reads, max_reads, subsamples)
        want = [O.effective_number_of_subsamples(int(r), min_reads, max_reads, subsamples) for r in R]
        assert effN.tolist() == want
        assert evaluated.tolist() == [i for i, e in enumerate(want) if e]
        assert np.diff(draw_offsets).tolist() == [e for e in want if e]
        assert all(int(k_of[i]) == min(int(R[i]) - 1, max_reads) for i in evaluated.tolist())
