"""
Consumers of the per-window LLRs: window binning, per-bin mean / standard error and crossover
detection.  CPU part: the oracle against vectors made by running the unmodified reference
(tests/golden/make_golden_consumers.py) and the product's host logic (bin bookkeeping, crossover scan);
GPU part: the per-bin statistics kernel through the C-ABI.
"""
import contextlib, io
import numpy as np
import pytest

import ldpgta_oracle as O
from conftest import load_golden, rel_err


@pytest.fixture(scope='module')
def golden():
    return load_golden('consumers.p.bz2')


def close(a, b, tol):
    if a is None or b is None:
        return a is b
    return rel_err(a, b) <= tol


def check_crossovers(fn, rec, tol=1e-12):
    for (z, look), (v1, v2) in rec['crossovers'].items():
        means, variances = rec['mean_var']
        got = fn(list(rec['LLRs']), means, variances, z_score=z, lookahead=look, recover=False)
        assert list(got) == list(v1) and all(close(got[k], v1[k], tol) for k in v1)
        if isinstance(v2, dict):
            with contextlib.redirect_stdout(io.StringIO()):
                got = fn(list(rec['LLRs']), means, variances, z_score=z, lookahead=look, recover=True)
            assert list(got) == list(v2) and all(close(got[k], v2[k], tol) for k in v2)
        else:                                   # the reference raises (empty candidate range); so must the restatements
            with pytest.raises(ValueError), contextlib.redirect_stdout(io.StringIO()):
                fn(list(rec['LLRs']), means, variances, z_score=z, lookahead=look, recover=True)


def test_oracle_bins_and_binning(golden):
    for name, rec in golden.items():
        info = {'chr_id': rec['chr_id']}
        for nb, bins in rec['bins'].items():
            got = O.bin_genomic_windows(list(rec['LLRs']), rec['chr_id'], nb)
            assert got == bins and list(got) == list(bins), (name, nb)
            for variant in ('plot', 'roc'):
                X, Y, E = O.binning(rec['LLRs'], info, nb, rounding=variant)
                gX, gY, gE = rec['binning_' + variant][nb]
                assert X == gX
                assert all(close(a, b, 1e-13) for a, b in zip(Y, gY)) and all(close(a, b, 1e-13) for a, b in zip(E, gE)), (name, nb)


def test_oracle_crossovers(golden):
    assert any(isinstance(v2, dict) and len(v2) > len(v1) for rec in golden.values() for v1, v2 in rec['crossovers'].values())
    assert any(v2 == 'ValueError' for rec in golden.values() for v1, v2 in rec['crossovers'].values())
    for rec in golden.values():
        check_crossovers(O.detect_crossovers, rec)


def test_product_bins_and_crossovers_host(golden):
    from ldpgta_b200 import llr_bins as B
    for name, rec in golden.items():
        for nb, bins in rec['bins'].items():
            got = B.bin_genomic_windows(list(rec['LLRs']), rec['chr_id'], nb)
            assert got == bins and list(got) == list(bins), (name, nb)
        check_crossovers(lambda *a, recover=False, **k: (B.detect_crossovers_v2 if recover else B.detect_crossovers)(*a, **k), rec)
    assert B.chr_length('chr21') == 46709983 and B.LLR(0.0, 0.5) == 1.23456789 * -1 and B.LLR(0, 0) == 0
    with pytest.raises(KeyError):
        B.chr_length('chr99')


@pytest.mark.gpu
def test_binning_on_device(golden):
    from ldpgta_b200 import llr_bins as B
    worst = 0.0
    for name, rec in golden.items():
        info = {'chr_id': rec['chr_id']}
        for nb in rec['bins']:
            X, Y, E = B.binning(rec['LLRs'], info, nb)
            for variant in ('plot', 'roc'):
                gX, gY, gE = rec['binning_' + variant][nb]
                assert X == gX and [y is None for y in Y] == [y is None for y in gY]
                for a, b in list(zip(Y, gY)) + list(zip(E, gE)):
                    if a is not None:
                        worst = max(worst, abs(a - b) / max(abs(b), 1e-300) if b else abs(a))
    assert worst <= 1e-9, worst               # tolerance of the chromosome-level LLR statistics (north_star)


@pytest.mark.gpu
def test_binning_pairs_keeps_llrs_resident(golden_inputs):
    """ likelihoods -> LLRs of the five pairs on the device -> per-bin statistics, without the LLRs leaving the device;
    equal to binning() fed with host-side LLRs of each pair. """
    from conftest import chrom_case
    from ldpgta_b200 import llr_bins as B
    from ldpgta_b200.engine import LLR_PAIRS
    args, out = chrom_case('homog_m3o', golden_inputs)
    lik = {w: [O.likelihoods_tuple(*l) for l in ls] for w, ls in out['likelihoods'].items()}
    info = {'chr_id': 'chr21'}
    res = B.binning_pairs(lik, info, 200)
    assert set(res) == set(LLR_PAIRS)
    for pair in LLR_PAIRS:
        llrs = {w: tuple(O.LLR(getattr(l, pair[0]), getattr(l, pair[1])) for l in ls) for w, ls in lik.items()}
        X, Y, E = O.binning(llrs, info, 200)
        gX, gY, gE = res[pair]
        assert gX == X
        assert all(close(a, b, 1e-9) for a, b in zip(gY, Y)) and all(close(a, b, 1e-9) for a, b in zip(gE, E))


@pytest.mark.gpu
def test_bin_stats_errors():
    from ldpgta_b200.engine import Engine
    from ldpgta_b200._native import LdpgtaError
    eng = Engine([8])
    off = np.array([0, 3, 6, 9])
    x = np.arange(9.0)
    with pytest.raises(LdpgtaError):
        eng.bin_stats(x, off, [[1, 1]])                 # empty bin
    with pytest.raises(LdpgtaError):
        eng.bin_stats(x, off, [[0, 4]])                 # beyond the last window
    with pytest.raises(LdpgtaError):
        eng.bin_stats(None, off, [[0, 3]])              # nothing resident yet
    got = eng.bin_stats(x, off, [[0, 3], [1, 2]])
    mean, std = O.mean_and_std_of_mean_of_rnd_var([(0., 1., 2.), (3., 4., 5.), (6., 7., 8.)])
    assert rel_err(got[0, 0, 0], mean) < 1e-15 and rel_err(got[0, 0, 1], std) < 1e-12
    assert got.shape == (1, 2, 2) and rel_err(got[0, 1, 0], 4.0) < 1e-15
