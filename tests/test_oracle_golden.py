"""Pins oracle/ldpgta_oracle.py to vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import hashlib

import pytest

import ldpgta_oracle as O
from conftest import CHROM_CONFIGS, chrom_case, load_golden, rel_err

LIK_TOL = 1e-14        # oracle vs reference likelihoods (different float summation order only)


def canon(model):
    out = {}
    for nblocks, classes in model.items():
        c = {}
        for w, members in classes.items():
            if isinstance(members, dict):
                c[w] = sorted((b0, tuple(sorted(v))) for b0, v in members.items())
            else:
                c[w] = sorted(members)
        out[nblocks] = sorted(c.items())
    return sorted(out.items())


def digest(models_n):
    return hashlib.sha256(repr({k: canon(v) for k, v in sorted(models_n.items())}).encode()).hexdigest()


@pytest.fixture(scope='module')
def models():
    return O.make_models(9)


def test_model_tables_match_reference_build(models):
    ref = load_golden('models_ref.p.bz2')
    for n, m in ref['models'].items():
        for scenario in ('MONOSOMY', 'DISOMY', 'SPH', 'BPH'):
            assert canon(models[n][scenario]) == canon(m[scenario]), (n, scenario)
    for n, d in ref['digests'].items():
        assert digest(models[n]) == d, n


def test_hap_dict_and_matches(golden_testmodule):
    g = golden_testmodule
    for grp in g['hap_tab']:
        hd, frac = O.build_hap_dict(g['obs_tab'], g['leg_tab'], g['hap_tab'][grp], g['number_of_haplotypes'][grp])
        assert frac == g['fraction_of_matches'][grp][grp]
        assert len(hd) > 900


def test_joint_counts_bit_exact(golden_testmodule):
    g = golden_testmodule
    hds = {grp: O.build_hap_dict(g['obs_tab'], g['leg_tab'], g['hap_tab'][grp], g['number_of_haplotypes'][grp])[0]
           for grp in g['hap_tab']}
    for rec in g['records']:
        for grp, hd in hds.items():
            masks = O.read_masks(hd, rec['alleles'])
            counts = O.joint_counts(masks)
            assert {s: c for s, c in enumerate(counts) if s} == rec['homog_counts_' + grp]
            if rec['n'] <= 6:
                assert counts == O.joint_counts_redundant(masks)


def test_likelihoods_three_model_classes(golden_testmodule, models):
    g = golden_testmodule
    obs, leg, hap, nh = g['obs_tab'], g['leg_tab'], g['hap_tab'], g['number_of_haplotypes']
    homog = {grp: O.Homogeneous(obs, leg, {grp: hap[grp]}, {grp: nh[grp]}, models) for grp in hap}
    recent = O.RecentAdmixture(obs, leg, hap, nh, models)
    distant = [O.DistantAdmixture(obs, leg, hap, nh, models, mk) for mk in g['makeups']]
    worst = 0.0
    for rec in g['records']:
        al = rec['alleles']
        for grp, m in homog.items():
            for mine, ref in ((m.likelihoods(al), rec['homog_general_' + grp]), (m.get_likelihoods(al), rec['homog_get_' + grp])):
                worst = max(worst, max(rel_err(a, b) for a, b in zip(mine, ref)))
        for mine, ref in ((recent.likelihoods(al), rec['recent_general']), (recent.get_likelihoods(al), rec['recent_get'])):
            worst = max(worst, max(rel_err(a, b) for a, b in zip(mine, ref)))
        for i, d in enumerate(distant):
            freq = d.effective(al)
            assert all(rel_err(freq[s], v) < 1e-15 for s, v in rec['distant%d_freq' % i].items())
            worst = max(worst, max(rel_err(a, b) for a, b in zip(d.likelihoods(al), rec['distant%d_general' % i])))
            if rec['n'] == 5:
                assert rec['distant%d_get_error' % i] == 'AttributeError'
                with pytest.raises(AttributeError):
                    d.get_likelihoods(al)
            else:
                worst = max(worst, max(rel_err(a, b) for a, b in zip(d.get_likelihoods(al), rec['distant%d_get' % i])))
    assert worst < LIK_TOL, worst


def test_exact_rational_agrees(golden_testmodule):
    g = golden_testmodule
    grp = 'group0'
    hd, _ = O.build_hap_dict(g['obs_tab'], g['leg_tab'], g['hap_tab'][grp], g['number_of_haplotypes'][grp])
    for rec in g['records']:
        if rec['n'] > 8:
            continue
        F = O.joint_counts(O.read_masks(hd, rec['alleles']))
        exact = O.likelihoods_exact(F, g['number_of_haplotypes'][grp], rec['n'])
        for e, r in zip(exact, rec['homog_general_' + grp]):
            assert rel_err(float(e), r) < 1e-14


@pytest.mark.parametrize('name', CHROM_CONFIGS)
def test_bootstrap_replay_and_statistics(name, golden_inputs, models):
    args, out = chrom_case(name, golden_inputs)
    cfg = args.pop('cfg')
    obs = O.SupportingDictionaries(args['obs_tab'], args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'],
                                   args['ancestral_makeup'], cfg['min_HF'])
    assert dict(obs.score) == out['score']                       # f1: read scores, exact integers
    lik, windows, matches = O.bootstrap(args['obs_tab'], args['leg_tab'], args['hap_tab'], args['number_of_haplotypes'],
                                        args['ancestral_makeup'], models, cfg['window_size'], cfg['subsamples'],
                                        cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'],
                                        cfg['min_HF'], replay=out['draws'])
    assert {w: frozenset(r) for w, r in windows.items()} == {w: frozenset(r) for w, r in out['windows'].items()}
    assert list(windows) == list(out['windows'])
    assert matches == out['fraction_of_matches']
    assert list(lik) == list(out['likelihoods'])
    worst = 0.0
    for w, ls in lik.items():
        assert len(ls) == len(out['likelihoods'][w])
        for mine, ref in zip(ls, out['likelihoods'][w]):
            worst = max(worst, max(rel_err(a, b) for a, b in zip(mine, ref)))
    assert worst < LIK_TOL, worst
    # a10: statistics on the reference's own likelihoods must reproduce its numbers
    ref_lik = {w: tuple(O.likelihoods_tuple(*l) for l in ls) for w, ls in out['likelihoods'].items()}
    st = O.statistics(ref_lik, out['windows'])
    gs = out['statistics']
    for key in ('num_of_windows', 'reads_mean', 'reads_std', 'window_size_mean', 'window_size_std'):
        assert st[key] == gs[key]
    for pair in O.LLR_PAIRS:
        assert st['LLRs_per_chromosome'][pair] == gs['LLRs_per_chromosome'][pair]
        assert st['LLRs_per_genomic_window'][pair] == gs['LLRs_per_genomic_window'][pair]


def test_effective_number_of_subsamples():
    assert O.effective_number_of_subsamples(5, 6, 4, 32) == 0
    assert O.effective_number_of_subsamples(6, 6, 4, 32) == 15
    assert O.effective_number_of_subsamples(75, 6, 4, 32) == 32
    assert O.effective_number_of_subsamples(4, 3, 4, 32) == 4
    assert O.effective_number_of_subsamples(3, 3, 4, 2) == 2


def test_llr_sentinels():
    assert O.LLR(0.0, 0.5) == -1.23456789
    assert O.LLR(0.5, 0.0) == 1.23456789
    assert O.LLR(0.0, 0.0) == 0
    assert O.LLR(0.25, 0.5) == pytest.approx(-0.6931471805599453, rel=1e-15)


def test_chr21_full_panel_bootstrap_replay():
    """ configs[0] of BASELINE.json: chr21-sized disomy sample at 0.05x on a 5008-haplotype panel through the UNMODIFIED
    reference bootstrap() / statistics() (tests/golden/make_golden_large.py, SNP density reduced to keep the fixture
    small; the panel is regenerated from its seed).  The oracle replays the recorded draws. """
    import hashlib
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden_large as G
    out = load_golden('chr21_5008.p.bz2')
    leg, rows = G.chr21_panel()
    assert hashlib.sha256(b''.join(r.to_bytes(626, 'little') for r in rows)).hexdigest() == out['panel_sha256']
    cfg = out['cfg']
    lik, windows, matches = O.bootstrap(out['obs_tab'], leg, {'ALL': rows}, {'ALL': G.CHR21_HAP}, {'ALL'}, O.make_models(4),
                                        cfg['window_size'], cfg['subsamples'], cfg['offset'], cfg['min_reads'], cfg['max_reads'],
                                        cfg['min_score'], cfg['min_HF'], replay=out['draws'])
    assert list(windows) == list(out['windows']) and matches == out['fraction_of_matches']
    assert {w: frozenset(r) for w, r in windows.items()} == {w: frozenset(r) for w, r in out['windows'].items()}
    assert list(lik) == list(out['likelihoods']) and len(lik) > 400
    worst = max(rel_err(a, b) for w in lik for mine, ref in zip(lik[w], out['likelihoods'][w]) for a, b in zip(mine, ref))
    assert worst < LIK_TOL, worst
    st = O.statistics({w: tuple(O.likelihoods_tuple(*l) for l in ls) for w, ls in out['likelihoods'].items()}, out['windows'])
    for pair in O.LLR_PAIRS:
        assert st['LLRs_per_chromosome'][pair] == out['statistics']['LLRs_per_chromosome'][pair]
