"""world_size-2 gloo test (CPU) of the N>1 host logic: sharding and the variable-length all-gather
that replicates a banded build."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from regridding_b200 import _parallel

        # a "public layout" sorted by input cell, split into input-row bands
        ncx, ncy = 9, 4
        rng = np.random.default_rng(0)
        ii = np.sort(rng.integers(0, ncx * ncy, 200))
        io = rng.integers(0, 50, 200)
        v = rng.random(200)
        lo, hi = _parallel.band_cells(ncx, ncy, rank, world)
        sel = (ii >= lo) & (ii < hi)
        parts = [_parallel.allgather_concat(torch.from_numpy(a[sel])) for a in (ii, io, v)]
        ok = all(np.array_equal(p.numpy(), a) for p, a in zip(parts, (ii, io, v)))
        # the list form gathers all three arrays of a band in one batch
        both = _parallel.allgather_concat([torch.from_numpy(a[sel]) for a in (ii, io, v)])
        ok = ok and all(np.array_equal(p.numpy(), a) for p, a in zip(both, (ii, io, v)))
        # frame sharding covers every frame exactly once
        f_lo, f_hi = _parallel.shard_range(11, rank, world)
        mine = torch.zeros(11, dtype=torch.int64)
        mine[f_lo:f_hi] = 1
        dist.all_reduce(mine)
        ok = ok and bool((mine == 1).all())
        # empty shard on one rank
        e = _parallel.allgather_concat(torch.arange(3 if rank == 0 else 0, dtype=torch.float64))
        ok = ok and e.tolist() == [0.0, 1.0, 2.0]
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_banded_allgather_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as manager:
        ret = manager.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
