"""World-size-2 gloo test (CPU) of the N>1 path: windows are dealt to ranks, each rank forms its
chromosome partials, one all_reduce adds them, and the result equals the reference's
mean_and_std_of_mean_of_rnd_var / fraction_of_negative_LLRs on the whole chromosome.  On the GPU the
partials come from ldpgta_window_stats; here they are formed with numpy from oracle LLRs, so the test
covers the sharding logic, the collective and ldpgta_finish_chromosome (host arithmetic in the .so)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import ldpgta_oracle as O
from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _chromosome(seed=5, n_windows=37):
    rng = np.random.default_rng(seed)
    counts = rng.integers(3, 12, size=n_windows)
    lik = {}
    for w, k in enumerate(counts):
        vals = rng.random((k, 4)) * 1e-3
        vals[rng.random(k) < 0.1, 2] = 0.0
        lik[(w * 100, w * 100 + 99)] = tuple(O.likelihoods_tuple(*v) for v in vals)
    reads = {w: tuple(range(int(rng.integers(6, 40)))) for w in lik}
    return lik, reads


def _partials(lik_items, n_cols):
    """ chrom_partials layout of ldpgta_window_stats: [0] sum of all LLRs, [1] windows, [2] windows with mean<0, [3+i] column sums. """
    out = np.zeros((5, n_cols + 3))
    for p, (i, j) in enumerate(O.LLR_PAIRS):
        for _w, ls in lik_items:
            x = [O.LLR(getattr(l, i), getattr(l, j)) for l in ls]
            out[p, 0] += sum(x)
            out[p, 1] += 1
            out[p, 2] += (sum(x) / len(x)) < 0
            out[p, 3:] += x[:n_cols]
    return out


def _worker(rank, world, port, q, n_windows=37):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import torch.distributed as dist
    from ldpgta_b200 import sharding
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    lik, reads = _chromosome(n_windows=n_windows)
    mine = sharding.shard_windows(reads, rank, world, min_reads=6, max_reads=4, subsamples=32, row_words=79)
    items = [(w, lik[w]) for w in mine]
    if items:
        n_cols_local = min(len(ls) for _, ls in items)
        first = len(items[0][1]) if mine[0] == next(iter(lik)) else 0
        res = sharding.allreduce_chromosome(_partials(items, n_cols_local), n_cols_local, first)
    else:                                                   # a rank without windows still joins the collectives
        res = sharding.allreduce_chromosome(None, None, 0)
    q.put((rank, [w for w in mine], res.tolist()))
    dist.destroy_process_group()


def test_two_rank_chromosome_statistics():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lik, reads = _chromosome()
    # the shards are contiguous, disjoint and cover the chromosome in order
    assert got[0][1] + got[1][1] == list(lik)
    assert got[0][1] and got[1][1]
    want = O.statistics(lik, reads)['LLRs_per_chromosome']
    for r in range(world):
        res = np.array(got[r][2])
        for p, pair in enumerate(O.LLR_PAIRS):
            assert abs(res[p, 0] - want[pair]['mean_of_mean']) < 1e-12
            assert abs(res[p, 1] - want[pair]['std_of_mean']) < 1e-12
            assert res[p, 2] == want[pair]['fraction_of_negative_LLRs']


def test_more_ranks_than_windows():
    """ world_size 3, one window: two ranks own nothing, pass the empty contribution, and all three get the statistics. """
    world, port = 3, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, 1)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lik, reads = _chromosome(n_windows=1)
    assert sorted(len(g[1]) for g in got) == [0, 0, 1]
    want = O.statistics(lik, reads)['LLRs_per_chromosome']
    for r in range(world):
        res = np.array(got[r][2])
        for p, pair in enumerate(O.LLR_PAIRS):
            assert abs(res[p, 0] - want[pair]['mean_of_mean']) < 1e-12
            assert res[p, 2] == want[pair]['fraction_of_negative_LLRs']


class _NoPeerEngine:
    """ stands for an Engine on a box where peer memory cannot be mapped: rank 1's buffer allocation fails """
    def __init__(self, rank):
        self.rank = rank

    def comm_create(self, rank, world, n):
        if self.rank == 1:
            raise RuntimeError('no peer access')
        return b'x' * 128

    def comm_connect(self, handles):
        raise AssertionError('must not be reached when a rank has no handle')


def _peer_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ldpgta_b200 import sharding
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    q.put((rank, sharding.connect_peers(_NoPeerEngine(rank), 1000)))
    dist.destroy_process_group()


def test_connect_peers_degrades_without_hanging():
    """ One rank cannot set up its exchange buffer: every rank still passes the collectives of connect_peers and all of
    them learn that the peer path is off (the callers then use all_reduce). """
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(0, False), (1, False)]


def test_partition_is_balanced_and_contiguous():
    sys.path.insert(0, ROOT)
    from ldpgta_b200 import sharding
    rng = np.random.default_rng(1)
    costs = rng.integers(0, 1000, size=500)
    for world in (1, 2, 3, 8):
        parts = sharding.partition_windows(costs, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(costs)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        loads = [costs[a:b].sum() for a, b in parts]
        assert max(loads) - min(loads) <= 2 * costs.max()
    assert sharding.partition_windows([0, 0, 0, 0], 2) == [(0, 2), (2, 4)]
    assert sharding.draw_cost(5, 6, 4, 32, 79) == 0
    assert sharding.draw_cost(75, 6, 4, 32, 79) == 32 * 15 * 79
    assert sharding.draw_cost(4, 3, 4, 32, 79) == 4 * 7 * 79
