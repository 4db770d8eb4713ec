"""
Multi-GPU sharding of the likelihood path: one process per GPU, windows dealt to ranks as
contiguous ranges balanced by work, no data-path collective.  The only exchange is the
per-chromosome LLR summary of statistics() (ANEUPLOIDY_TEST.py:65-76, 309-314) when one
chromosome is split across ranks: every rank contributes its chrom_partials
([5][n_cols + 3] float64, see ldpgta_window_stats) to ONE all_reduce(sum), and
ldpgta_finish_chromosome turns the sum into mean_of_mean / std_of_mean /
fraction_of_negative_LLRs.  torch.distributed is plumbing only (NCCL on GPUs, gloo in CPU tests).
"""
from math import comb

import numpy as np

from .engine import finish_chromosome


def draw_cost(n_reads_in_window, min_reads, max_reads, subsamples, row_words, n_groups=1):
    """ Algorithmic word-ops of one window: effN * (2^n - 1) * W * G. """
    R = n_reads_in_window
    if min_reads <= R and R > max_reads:
        effN = min(comb(R, max_reads), subsamples)
    elif min_reads <= R <= max_reads:
        effN = min(R, subsamples)
    else:
        return 0
    n = min(R - 1, max_reads)
    return effN * ((1 << n) - 1) * row_words * n_groups


def partition_windows(costs, world_size):
    """ Contiguous ranges [lo, hi) per rank with near-equal total cost (prefix-sum split). """
    costs = np.asarray(costs, dtype=np.float64)
    total = costs.sum()
    bounds = [0]
    if total <= 0:
        step = len(costs) / world_size
        return [(int(round(r * step)), int(round((r + 1) * step))) for r in range(world_size)]
    prefix = np.concatenate([[0.0], np.cumsum(costs)])
    for r in range(1, world_size):
        bounds.append(int(np.searchsorted(prefix, total * r / world_size, side='left')))
    bounds.append(len(costs))
    bounds = [min(max(b, bounds[i - 1] if i else 0), len(costs)) for i, b in enumerate(bounds)]
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_windows(genomic_windows, rank, world_size, min_reads, max_reads, subsamples, row_words, n_groups=1):
    """ The windows (keys of genomic_windows, in order) owned by `rank`. """
    keys = list(genomic_windows)
    costs = [draw_cost(len(genomic_windows[k]), min_reads, max_reads, subsamples, row_words, n_groups) for k in keys]
    lo, hi = partition_windows(costs, world_size)[rank]
    return keys[lo:hi]


def allreduce_chromosome(partials, n_cols_local, first_window_draws, group=None, device=None):
    """ Sum of the ranks' chromosome partials -> float64[5, 3] (mean_of_mean, std_of_mean,
    fraction_of_negative_LLRs).  n_cols is the smallest draw count over ALL windows of the
    chromosome (zip truncation, ANEUPLOIDY_TEST.py:73): ranks first agree on it (MIN), truncate
    their column sums, then add.  first_window_draws = len(A[0]) of the reference, taken from the
    rank that owns the chromosome's first window (others pass 0; MAX picks it).
    A rank that owns NO window of the chromosome (more ranks than windows, or the cost concentrated in
    a few windows) still takes part in both collectives: it passes partials=None and n_cols_local=None,
    contributing nothing to the minimum and zeros to the sum.  With no window anywhere the result is NaN. """
    import torch
    import torch.distributed as dist
    empty = partials is None or n_cols_local is None
    if not (dist.is_available() and dist.is_initialized()):
        if empty:
            return np.full((5, 3), np.nan)
        partials = np.asarray(partials, dtype=np.float64)
        return finish_chromosome(np.ascontiguousarray(partials[:, :n_cols_local + 3]), n_cols_local, first_window_draws)
    dev = device if device is not None else ('cuda' if dist.get_backend(group) == 'nccl' else 'cpu')
    meta = torch.tensor([-np.inf if empty else -float(n_cols_local), 0.0 if empty else float(first_window_draws)],
                        dtype=torch.float64, device=dev)
    dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=group)          # max(-n_cols) = -min(n_cols)
    if not np.isfinite(meta[0].item()):
        return np.full((5, 3), np.nan)                                 # no rank holds a window of this chromosome
    n_cols, first = int(-meta[0].item()), int(meta[1].item())
    mine = np.zeros((5, n_cols + 3)) if empty else np.ascontiguousarray(np.asarray(partials, dtype=np.float64)[:, :n_cols + 3])
    t = torch.from_numpy(mine).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return finish_chromosome(t.cpu().numpy(), n_cols, first)


def connect_peers(engine, n_doubles, group=None):
    """ Sets up the peer-memory exchange of ldpgta_allreduce_chrom_dev between the ranks of a torch.distributed group
    (one process per GPU): every rank allocates its buffer, the handles travel by all_gather_object (plumbing only), and
    each rank maps the peers' buffers.  Afterwards engine.allreduce_chrom_dev(ptr, n) sums device buffers across the
    ranks in one kernel over NVLink, with no collective library in the data path.
    Returns True when EVERY rank is connected.  A rank that fails (no peer access between two GPUs, IPC disabled in the
    container) still takes part in the collectives below, so nobody hangs; the callers then keep using all_reduce. """
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    try:
        mine = engine.comm_create(rank, world, n_doubles)
    except Exception:
        mine = None
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    ok = all(h is not None for h in handles)
    if ok:
        try:
            engine.comm_connect(handles)
        except Exception:
            ok = False
    dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # also: every rank has mapped every buffer before the first exchange
    return bool(flag.item())
