"""The real multi-process sharded builds (exchange-free band build; line-sharded with NCCL / peer-mapped memory) on 2 GPUs of one box; skipped on
single-GPU boxes (there the same kernels are covered by test_build2d_line_sharded_equals_full, which plays all
ranks on one device)."""

import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from regridding_b200 import _device, _parallel
        from tests import cases

        gi, go, _ = cases.case_2d("dist129")
        co = cases.perturb_like_reference(go, (-1, -2), 42)
        t = [torch.from_numpy(a).to(dev) for a in (*gi, *co)]
        full = _device.build_weights_2d(*t, device=dev)
        ok = True
        for exchange in ("band", "p2p", "nccl"):
            rep = _parallel.build_weights_2d_sharded(*t, replicate=True, device=dev, exchange=exchange)
            ok = ok and all(torch.equal(getattr(rep, k), getattr(full, k))
                            for k in ("indices_input", "indices_output", "values"))
            band = _parallel.build_weights_2d_sharded(*t, replicate=False, device=dev, exchange=exchange)
            lo, hi = _parallel.band_cells(gi[0].shape[0] - 1, gi[0].shape[1] - 1, rank, world)
            sel = (full.indices_input >= lo) & (full.indices_input < hi)
            ok = ok and torch.equal(band.values, full.values[sel]) and torch.equal(band.indices_output, full.indices_output[sel])
        # band build with buffers that are too small at first: every rank repeats the build with the learned sizes
        for key in list(_device._band_caps):
            _device._band_caps[key] = (1000, 500)
        rep = _parallel.build_weights_2d_sharded(*t, replicate=True, device=dev, exchange="band")
        ok = ok and torch.equal(rep.values, full.values) and torch.equal(rep.indices_input, full.indices_input)
        # arena growth: start with room for 1 fragment per cell -> every rank overflows, all grow together, walk again
        os.environ["REGRID_B200_ARENA_FRAGS_PER_CELL"] = "1"
        _parallel._arenas.clear()
        rep = _parallel.build_weights_2d_sharded(*t, replicate=True, device=dev, exchange="p2p")
        ok = ok and torch.equal(rep.values, full.values) and torch.equal(rep.indices_input, full.indices_input)
        grown = next(iter(_parallel._arenas.values())).capacity
        ok = ok and grown >= full.stats["fragments"] // world // 2
        os.environ.pop("REGRID_B200_ARENA_FRAGS_PER_CELL")
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_line_sharded_build_two_processes():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    port = _free_port()
    with mp.Manager() as manager:
        ret = manager.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
