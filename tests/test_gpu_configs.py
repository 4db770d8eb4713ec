"""Parity at the shapes of BASELINE.json's configs (the bench line covers configs[1]; the others are
parity cases): synthetic LD-structured read masks from bench.py's generator, draws planned like the
reference does, CUDA path through the C-ABI against the C oracle on samples, plus size-independent
properties at full size."""
import numpy as np
import pytest
import torch

import bench
import coracle as C
import ldpgta_b200 as L

pytestmark = pytest.mark.gpu
LIK_TOL = 1e-12


def max_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300))) if a.size else 0.0


def workload(n_windows, reads_mean, n_hap, max_reads, subsamples, seed):
    rng = np.random.default_rng(seed)
    reads = np.maximum(rng.poisson(reads_mean, size=n_windows), max_reads + 2).astype(np.int64)
    masks = bench.synth_masks(torch, torch.device('cuda', 0), reads, n_hap, seed)
    idx, win = bench.plan_draws(reads, subsamples, max_reads, seed + 1)
    return reads, masks, idx, win


def host_rows(masks, n_hap):
    return masks.cpu().numpy().view(np.uint64)[:, :(n_hap + 63) // 64].copy()


def test_config3_one_x_coverage_sixteen_reads():
    """ configs[2]: 1x coverage (about 1500 reads per window), 16 reads per draw: 2^16 - 1 subsets. """
    reads, masks, idx, _ = workload(n_windows=6, reads_mean=1500, n_hap=5008, max_reads=16, subsamples=4, seed=31)
    n_draws = idx.shape[0]
    with L.Engine([5008]) as eng:
        eng.attach_read_masks_dev(0, masks.data_ptr(), masks.shape[0], keepalive=masks)
        off = np.arange(0, 16 * n_draws + 1, 16, dtype=np.int64)
        got = eng.likelihoods_batch(L.HOMOGENEOUS, idx.ravel(), off)
        again = eng.likelihoods_batch(L.HOMOGENEOUS, idx.ravel(), off)
        assert np.array_equal(got, again)
        rows = host_rows(masks, 5008)
        k = 4
        want = C.batch([rows], [5008], idx[:k].ravel(), off[:k + 1], 0)
        assert max_rel(got[:k], want) < LIK_TOL
        cnt = eng.joint_counts(idx[0])
        assert np.array_equal(cnt[0], C.joint_counts(rows[idx[0]]))
        assert np.all(got >= 0) and np.all(got <= 1) and np.all(np.isfinite(got))


def test_config4_recent_admixture_two_population_panel():
    """ configs[3]: two populations of 2504 haplotypes, admixed embryo at 0.1x, defaults (4 reads per draw). """
    reads, m0, idx, _ = workload(n_windows=400, reads_mean=150, n_hap=2504, max_reads=4, subsamples=32, seed=41)
    m1 = bench.synth_masks(torch, torch.device('cuda', 0), reads, 2504, 43)
    n_draws = idx.shape[0]
    off = np.arange(0, 4 * n_draws + 1, 4, dtype=np.int64)
    with L.Engine([2504, 2504]) as eng:
        eng.attach_read_masks_dev(0, m0.data_ptr(), m0.shape[0], keepalive=m0)
        eng.attach_read_masks_dev(1, m1.data_ptr(), m1.shape[0], keepalive=m1)
        got = eng.likelihoods_batch(L.RECENT_ADMIXTURE, idx.ravel(), off)
        sample = np.random.default_rng(1).choice(n_draws, size=1500, replace=False)
        rows = [host_rows(m0, 2504), host_rows(m1, 2504)]
        want = C.batch(rows, [2504, 2504], idx[sample].ravel(), np.arange(0, 4 * 1500 + 1, 4, dtype=np.int64), 1)
        assert max_rel(got[sample], want) < LIK_TOL
        swapped = None
    with L.Engine([2504, 2504]) as eng:              # the model is symmetric in the two populations
        eng.attach_read_masks_dev(0, m1.data_ptr(), m1.shape[0], keepalive=m1)
        eng.attach_read_masks_dev(1, m0.data_ptr(), m0.shape[0], keepalive=m0)
        swapped = eng.likelihoods_batch(L.RECENT_ADMIXTURE, idx.ravel(), off)
    assert max_rel(got, swapped) < 1e-13


def test_config5_thousand_subsamples_batch_of_embryos():
    """ configs[4] shape: several embryos that share window structure, 1000 draws per window; every embryo is an
    independent context on the same GPU; sampled oracle check and determinism. """
    n_hap = 5008
    for embryo in range(3):
        reads, masks, idx, win = workload(n_windows=40, reads_mean=150, n_hap=n_hap, max_reads=4, subsamples=1000, seed=50 + embryo)
        n_draws = idx.shape[0]
        assert n_draws == 40 * 1000
        off = np.arange(0, 4 * n_draws + 1, 4, dtype=np.int64)
        with L.Engine([n_hap]) as eng:
            eng.attach_read_masks_dev(0, masks.data_ptr(), masks.shape[0], keepalive=masks)
            got = eng.likelihoods_batch(L.HOMOGENEOUS, idx.ravel(), off)
            sample = np.random.default_rng(embryo).choice(n_draws, size=800, replace=False)
            want = C.batch([host_rows(masks, n_hap)], [n_hap], idx[sample].ravel(), np.arange(0, 4 * 800 + 1, 4, dtype=np.int64), 0)
            assert max_rel(got[sample], want) < LIK_TOL
            # per-window statistics over 1000 draws on the device
            woff = np.arange(0, n_draws + 1, 1000, dtype=np.int64)
            llr, mean_var, partials = eng.window_stats(got, woff, 1000)
            chrom = L.finish_chromosome(partials, 1000, 1000)
            x = llr[0].reshape(40, 1000)
            assert np.allclose(mean_var[0, :, 0], x.mean(axis=1), rtol=0, atol=1e-12)
            assert np.allclose(mean_var[0, :, 1], x.var(axis=1, ddof=1), rtol=1e-10, atol=1e-14)
            assert abs(chrom[0, 0] - x.mean()) < 1e-12
