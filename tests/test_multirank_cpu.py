"""The N>1 path on CPU: two gloo ranks shard an event range, accumulate their integer band words
and all-reduce them; the sum must be bit-identical to the single-rank accumulator."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference_lux_run03.npz")


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from nest_b200 import capi, shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    g = np.load(GOLD)
    s1, s2 = g["band_sig"]
    ana = capi.analysis(0)
    lo, hi = shard.event_range(rank, world, s1.size)
    words = torch.from_numpy(shard.band_words_from_signals(ana, 2, s1[lo:hi], s2[lo:hi]))
    dist.all_reduce(words)                 # int64 sum == sum of the u64 words
    if rank == 0:
        q.put(words.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_band_allreduce_is_exact():
    sys.path.insert(0, ROOT)
    from nest_b200 import capi, shard
    g = np.load(GOLD)
    s1, s2 = g["band_sig"]
    ana = capi.analysis(0)
    full = shard.band_words_from_signals(ana, 2, s1, s2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got, full)
    band = shard.band_from_words(ana, got).finalize()
    assert np.isfinite(band[:, 2]).sum() > 60


def _sweep_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from nest_b200 import shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n_points = 250
    mine = shard.sweep_points(rank, world, n_points)
    # every rank fills the slots of its own points; a sum over ranks must then touch every slot exactly once
    owner = torch.zeros(n_points, dtype=torch.int64)
    owner[mine] = 1
    which = torch.full((n_points,), 0, dtype=torch.int64)
    which[mine] = rank
    dist.all_reduce(owner)
    dist.all_reduce(which)
    if rank == 0:
        q.put((owner.numpy().copy(), which.numpy().copy(), len(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sweep_points_cover_the_grid_once():
    """config 4 at N > 1: grid points are dealt round-robin, each to exactly one rank (no collective on the data path)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    owner, which, n0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert (owner == 1).all() and n0 == 125
    assert np.array_equal(which, np.arange(250) % 2)
