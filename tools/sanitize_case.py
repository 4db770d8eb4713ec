#!/usr/bin/env python3
"""Small cases of every kernel for compute-sanitizer (one tool per gpurun call)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ldpgta_b200 as L

rng = np.random.default_rng(0)
def rows(n, nh):
    w = (nh + 63) // 64
    r = rng.integers(0, 2**63, size=(n, w), dtype=np.uint64) | rng.integers(0, 2**63, size=(n, w), dtype=np.uint64)
    t = nh - 64 * (w - 1)
    if t < 64: r[:, -1] &= np.uint64((1 << t) - 1)
    return r
with L.Engine([5008]) as e:
    e.upload_read_masks(0, rows(64, 5008))
    for n, nd in ((2, 9), (4, 37), (5, 3), (8, 3), (12, 2)):
        idx = np.stack([rng.choice(64, size=n, replace=False) for _ in range(nd)]).astype(np.int32).ravel()
        out = e.likelihoods_batch(L.HOMOGENEOUS, idx, np.arange(0, n * nd + 1, n))
        assert np.all(np.isfinite(out))
    e.joint_counts(np.arange(4, dtype=np.int32))
with L.Engine([331, 200]) as e:
    for g, nh in enumerate((331, 200)):
        e.upload_allele_rows(g, rows(50, nh), np.array([i % 2 for i in range(50)], dtype=np.uint8))
    off = np.arange(0, 41, 2); al = rng.integers(0, 50, size=40).astype(np.int32)
    e.build_read_masks(al, off)
    print('scores', e.read_scores(al, off, [0.5 / 331, 0.5 / 200], 0.5 / 200, 0.05)[:5])
    for kind, prop in ((L.RECENT_ADMIXTURE, None), (L.DISTANT_ADMIXTURE, [0.4, 0.6])):
        for n, nd in ((3, 5), (4, 5), (6, 2), (9, 2)):
            idx = np.stack([rng.choice(20, size=n, replace=False) for _ in range(nd)]).astype(np.int32).ravel()
            e.likelihoods_batch(kind, idx, np.arange(0, n * nd + 1, n), proportions=prop)
    lik = rng.random((40, 4))
    e.window_stats(lik, np.arange(0, 41, 8), 8)
    e.get_likelihoods(L.RECENT_ADMIXTURE, [[0, 1], [2], [3, 4, 5]])
print('sanitize case ok')
