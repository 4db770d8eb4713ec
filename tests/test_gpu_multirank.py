"""
The N > 1 path on real GPUs (needs two; skipped on a single-GPU box): one process per GPU, the windows of two
chromosomes dealt to the ranks by cost (sharding.partition_windows), every rank running the device-resident pipeline on
its windows, the segment partial sums added across the GPUs -- once with torch.distributed all_reduce (NCCL) and once with
the library's own peer-memory kernel (ldpgta_allreduce_chrom_dev over NVLink) -- and finished on the device.  The sharded
statistics must equal the single-GPU statistics of the same draws to 1e-12, and the two exchanges must agree bit for bit
across ranks.
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _sample(seed=21, n_windows=64, n_hap=331):
    rng = np.random.default_rng(seed)
    R = rng.integers(8, 30, size=n_windows)
    words = (n_hap + 63) // 64
    bits = rng.random((int(R.sum()), words * 64)) < 0.7
    bits[:, n_hap:] = False
    rows = (bits.reshape(-1, words, 64).astype(np.uint64) << np.arange(64, dtype=np.uint64)).sum(axis=2, dtype=np.uint64)
    woff = np.concatenate([[0], np.cumsum(R)]).astype(np.int64)
    per = rng.integers(5, 12, size=n_windows)
    doff = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
    seg = np.array([0, 27, n_windows], dtype=np.int64)
    return rows, woff, doff, per, seg, n_hap


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import ldpgta_b200 as L
    from ldpgta_b200 import sharding
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    rows, woff, doff, per, seg, n_hap = _sample()
    n = 4
    with L.Engine([n_hap], device=rank) as eng:
        eng.upload_read_masks(0, rows)
        # the whole sample on this GPU: reference statistics and the draws every shard replays
        eng.set_windows(woff, None, doff, n, segment_offsets=seg)
        whole = eng.bootstrap_stats(L.HOMOGENEOUS, 5, want=('draws', 'stats', 'partials'))
        draws = whole['draws'].reshape(-1, n)
        # this rank's windows
        costs = [sharding.draw_cost(int(woff[w + 1] - woff[w]), 6, n, int(per[w]), 6) for w in range(len(per))]
        lo, hi = sharding.partition_windows(costs, world)[rank]
        cols = np.array([per[seg[s]:seg[s + 1]].min() for s in range(2)], dtype=np.int32)
        first = np.array([per[seg[s]] for s in range(2)], dtype=np.int32)
        info = eng.set_windows(woff[lo:hi + 1], None, doff[lo:hi + 1] - doff[lo], n, segment_offsets=np.clip(seg, lo, hi) - lo,
                               segment_cols=cols, segment_first=first)
        mine = eng.replay_stats(L.HOMOGENEOUS, draws[doff[lo]:doff[hi]], want=('partials',))['partials']
        stream = torch.cuda.ExternalStream(eng.stream, device=torch.device('cuda', rank))
        out = {}
        for how in ('nccl', 'peer'):
            t = torch.from_numpy(mine.copy()).cuda(rank)
            stats = torch.zeros((2, 5, 3), dtype=torch.float64, device='cuda:%d' % rank)
            with torch.cuda.stream(stream):
                if how == 'nccl':
                    dist.all_reduce(t)
                else:
                    if 'connected' not in out:
                        assert sharding.connect_peers(eng, t.numel())
                        out['connected'] = True
                    for _ in range(3):                                   # repeated exchanges reuse the two buffers
                        t.copy_(torch.from_numpy(mine))
                        eng.allreduce_chrom_dev(t.data_ptr(), t.numel())
                eng.finish_segments_dev(t.data_ptr(), stats.data_ptr())
            eng.synchronize()
            torch.cuda.synchronize()
            out[how] = (t.cpu().numpy(), stats.cpu().numpy())
        dist.barrier()
        q.put((rank, (lo, hi), whole['stats'], whole['partials'], out['nccl'], out['peer']))
    dist.destroy_process_group()


def test_two_gpus_sharded_statistics_equal_single_gpu():
    torch = pytest.importorskip('torch')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run under gpurun --gpus 2)')
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=300) for _ in range(world)), key=lambda g: g[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0][1][1] == got[1][1][0] and got[0][1][0] == 0                 # contiguous shards
    for rank, _range, whole_stats, whole_part, (n_sum, n_stats), (p_sum, p_stats) in got:
        assert np.allclose(n_sum, whole_part, rtol=1e-13, atol=1e-12)
        assert np.max(np.abs(n_stats - whole_stats)) < 1e-12
        assert np.max(np.abs(p_stats - whole_stats)) < 1e-12
        assert np.allclose(p_sum, n_sum, rtol=1e-15, atol=1e-15)
    # the peer exchange adds in rank order on every rank: bit-identical sums everywhere
    assert np.array_equal(got[0][5][0], got[1][5][0]) and np.array_equal(got[0][5][1], got[1][5][1])
    assert np.array_equal(got[0][2], got[1][2])
