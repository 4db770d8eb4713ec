"""Pins the C checker (oracle/oracle_c.c) to the reference's golden vectors and to the CPython oracle. CPU only."""
import random

import numpy as np
import pytest

import coracle as C
import ldpgta_oracle as O
from conftest import rel_err


def _pad(rows_a, rows_b):
    W = max(rows_a.shape[1], rows_b.shape[1])
    out = []
    for r in (rows_a, rows_b):
        p = np.zeros((r.shape[0], W), dtype=np.uint64)
        p[:, :r.shape[1]] = r
        out.append(p)
    return out


def test_counts_and_likelihoods_vs_golden(golden_testmodule):
    g = golden_testmodule
    nh = g['number_of_haplotypes']
    hds = {grp: O.build_hap_dict(g['obs_tab'], g['leg_tab'], g['hap_tab'][grp], nh[grp])[0] for grp in g['hap_tab']}
    worst = 0.0
    for rec in g['records']:
        n = rec['n']
        rows = {grp: C.pack_rows(O.read_masks(hd, rec['alleles']), nh[grp]) for grp, hd in hds.items()}
        idx, off = np.arange(n, dtype=np.int32), np.array([0, n], dtype=np.int64)
        for grp in rows:
            cnt = C.joint_counts(rows[grp])
            assert {s: int(c) for s, c in enumerate(cnt) if s} == rec['homog_counts_' + grp]
            got = C.batch([rows[grp]], [nh[grp]], idx, off, 0)[0]
            worst = max(worst, max(rel_err(a, b) for a, b in zip(got, rec['homog_get_' + grp])))
        got = C.batch([rows['group0'], rows['group1']], [nh['group0'], nh['group1']], idx, off, 1)[0]
        worst = max(worst, max(rel_err(a, b) for a, b in zip(got, rec['recent_get'])))
        for i, mk in enumerate(g['makeups']):
            if n == 5:
                ref = rec['distant%d_general' % i]
            else:
                ref = rec['distant%d_get' % i]
            groups = list(mk)
            got = C.batch([rows[x] for x in groups], [nh[x] for x in groups], idx, off, 2, [mk[x] for x in groups])[0]
            worst = max(worst, max(rel_err(a, b) for a, b in zip(got, ref)))
    assert worst < 1e-14, worst


@pytest.mark.parametrize('n', [10, 12])
def test_large_n_vs_python_oracle(n):
    rng = random.Random(100 + n)
    N = 331
    founders = [rng.getrandbits(N) for _ in range(6)]
    masks = [founders[rng.randrange(6)] | rng.getrandbits(N) for _ in range(n)]   # dense, correlated rows
    F = O.joint_counts(masks)
    rows = C.pack_rows(masks, N)
    assert [0] + F[1:] == C.joint_counts(rows).tolist()
    exact = O.likelihoods_exact(F, N, n)
    got = C.batch([rows], [N], np.arange(n, dtype=np.int32), np.array([0, n], dtype=np.int64), 0)[0]
    for a, b in zip(got, exact):
        assert rel_err(a, float(b)) < 1e-15


def test_llr():
    for y, x in ((0.0, 0.5), (0.5, 0.0), (0.0, 0.0), (0.25, 0.5), (3e-7, 1e-9)):
        assert C.lib().oracle_llr(y, x) == pytest.approx(O.LLR(y, x), rel=1e-15, abs=0)


def large_draw_rows(rec, inputs):
    """ packed read masks of one large_draws record, per panel group (hap_dict of the oracle = build_hap_dict) """
    obs, leg = inputs['obs_tabs']['disomy'], inputs['leg_tab']
    rows = {}
    for g in ('G0', 'G1'):
        hd, _ = O.build_hap_dict(obs, leg, inputs['hap_tab'][g], inputs['number_of_haplotypes'][g])
        rows[g] = C.pack_rows(O.read_masks(hd, rec['alleles']), inputs['number_of_haplotypes'][g])
    return rows


def test_large_draws_vs_reference_build(golden_inputs):
    """ 10..16 (..18 when generated) reads per draw: the C checker against the UNMODIFIED reference run with
    MAKE_STATISTICAL_MODEL.BUILD tables (tests/golden/make_golden_large.py): joint-frequency tables bit-exact (sha256),
    likelihoods of all three model classes to 1e-14. """
    import hashlib
    from conftest import load_golden
    g = load_golden('large_draws.p.bz2')
    nh = golden_inputs['number_of_haplotypes']
    assert set(range(10, 19)) <= set(g['records'])
    worst = 0.0
    for n, rec in sorted(g['records'].items()):
        rows = large_draw_rows(rec, golden_inputs)
        cnt = C.joint_counts(rows['G0'])
        assert hashlib.sha256(cnt.astype('<u4').tobytes()).hexdigest() == rec['counts_sha256_G0'], n
        assert int(cnt[-1]) == rec['full_set_count_G0']
        if rec['counts_G0'] is not None:
            assert {s: int(c) for s, c in enumerate(cnt) if s} == rec['counts_G0']
        idx, off = np.arange(n, dtype=np.int32), np.array([0, n], dtype=np.int64)
        both = _pad(rows['G0'], rows['G1'])
        got = {'homog_G0': C.batch([rows['G0']], [nh['G0']], idx, off, 0)[0],
               'recent': C.batch(both, [nh['G0'], nh['G1']], idx, off, 1)[0],
               'distant': C.batch(both, [nh['G0'], nh['G1']], idx, off, 2, [rec['makeup']['G0'], rec['makeup']['G1']])[0]}
        for key, vals in got.items():
            worst = max(worst, max(rel_err(a, b) for a, b in zip(vals, rec[key])))
    assert worst < 1e-14, worst
