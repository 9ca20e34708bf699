"""CPU test of the multi-process host logic with gloo, world_size 2: rank-local scene seeding and the
max-over-ranks / sum-over-ranks aggregation bench.py uses for N > 1."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zeno_b200 import dist_util, scenes
    pos, vel, dx = scenes.dam_break_points(16, seed=1 + rank)
    ms, units = dist_util.aggregate(10.0 + rank, float(pos.shape[0]))
    slabs = dist_util.slab_bounds(64, world)
    if rank == 0:
        out.put((ms, units, slabs, pos.shape[0]))
    dist.destroy_process_group()


def test_two_rank_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, units, slabs, n = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == pytest.approx(11.0)       # max over ranks
    assert units == pytest.approx(2 * n)   # sum over ranks
    assert slabs == [(0, 32), (32, 64)]


def _slab_worker(rank, world, port, out):
    """host side of the slab decomposition: every rank derives the same balanced slabs from the all-reduced per-layer
    histogram, keeps the points of its slab, and the kept sets partition the scene"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from zeno_b200 import dist_util, scenes
    pos, vel, dx = scenes.dam_break_points(64, seed=3, side=32)
    # each rank starts from an arbitrary half of the points (what a distributed loader would hand it)
    mine = pos[rank::world]
    lx = dist_util.leaf_layer(mine, dx)
    hist = torch.zeros(16, dtype=torch.int64)
    hist.index_add_(0, torch.from_numpy(lx.astype(np.int64)), torch.ones(lx.shape[0], dtype=torch.int64))
    dist.all_reduce(hist)
    bounds = dist_util.balanced_slabs(hist.numpy(), world)
    lo, hi = bounds[rank]
    allx = dist_util.leaf_layer(pos, dx)
    kept = int(((allx >= lo) & (allx < hi)).sum())
    t = torch.tensor([kept], dtype=torch.int64)
    dist.all_reduce(t)
    if rank == 0:
        out.put((bounds, int(t.item()), pos.shape[0], hist.numpy().tolist()))
    dist.destroy_process_group()


def test_two_rank_balanced_slabs():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    bounds, kept, n, hist = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert kept == n                                     # the slabs partition the particles
    assert bounds[0][1] == bounds[1][0] and bounds[0][0] == 0 and bounds[1][1] == 4   # 32 voxels = 4 leaf layers
    assert all(hi - lo >= 2 for lo, hi in bounds)        # at least two layers per rank (flipb200_dd_set_slab)
    assert sum(hist) == n
