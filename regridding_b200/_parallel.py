"""
Multi-GPU sharding (one process per GPU, ``torch.distributed`` over NCCL/NVLink).

The path shards without any data-path collective:
  * independent orthogonal slices (spectra of config 2, frames of configs 3/4) are split
    into contiguous rank ranges (``shard_range``);
  * one large 2D build is split into contiguous INPUT-ROW bands: the public layout is
    sorted by input cell first (regridding/_weights/_weights_arrays.py:54-59), so the
    per-rank results concatenate in rank order with no re-sort.
The only collective is the optional all-gather of the band triplets when the caller
wants the full matrix replicated on every rank (e.g. before a frame-sharded apply).
A banded build still verifies every sweep line end to end (that is what keeps it
bit-identical to the full build), but the emit walk, the bucket sort and the merge only
touch the segments / fragments of the rank's own band.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from . import _device


def world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) share of n items; the first n % world ranks get one more."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def band_cells(ncx: int, ncy: int, rank: int, world_size: int) -> tuple[int, int]:
    """Flat input-cell range [lo, hi) of this rank's band of input rows."""
    r0, r1 = shard_range(ncx, rank, world_size)
    return r0 * ncy, r1 * ncy


def allgather_concat(tensors, group=None):
    """Variable-length all-gather along dim 0 of one tensor or of a list of tensors that share their
    dim-0 length: every rank gets ``cat([t_0, ..., t_{W-1}])`` of each.  One small collective for the
    lengths, then point-to-point copies of every band straight into its place in the result (one
    NCCL group, no padding, no concatenation pass); works under gloo on CPU too."""
    single = isinstance(tensors, torch.Tensor)
    ts = [tensors] if single else list(tensors)
    rank, W = world(group)
    if W == 1:
        return tensors
    dev = ts[0].device
    n = torch.tensor([ts[0].shape[0]], dtype=torch.int64, device=dev)
    counts = torch.empty(W, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts_h = counts.cpu().tolist()
    offs = [0]
    for c in counts_h:
        offs.append(offs[-1] + c)
    outs, ops = [], []
    for t in ts:
        t = t.contiguous()
        out = torch.empty((offs[-1],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        out[offs[rank]:offs[rank + 1]] = t
        for r in range(W):
            if r == rank:
                continue
            peer = r if group is None else dist.get_global_rank(group, r)
            if counts_h[rank]:
                ops.append(dist.P2POp(dist.isend, t, peer, group=group))
            if counts_h[r]:
                ops.append(dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], peer, group=group))
        outs.append(out)
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return outs[0] if single else outs


def build_weights_2d_banded(x_in, y_in, x_out, y_out, weights_input=None, replicate: bool = True,
                            group=None, device=None) -> _device.DeviceWeights:
    """Each rank builds the weights of its band of input rows; with ``replicate`` the bands
    are concatenated on every rank through an all-gather (the result then equals the
    single-GPU build bit for bit: the same segments are walked from the same states)."""
    rank, W = world(group)
    nxi, nyi = x_in.shape
    band = band_cells(nxi - 1, nyi - 1, rank, W)
    dw = _device.build_weights_2d(x_in, y_in, x_out, y_out, weights_input, cell_band=band, device=device)
    if not replicate or W == 1:
        return dw
    ii, io, v = allgather_concat([dw.indices_input, dw.indices_output, dw.values], group)
    out = _device.DeviceWeights(ii, io, v, dw.n_in, dw.n_out)
    out.stats = dw.stats
    return out
