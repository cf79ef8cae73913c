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

``build_weights_2d_sharded`` is the strong-scaling build: the sweep LINES are dealt out across
the ranks (every rank walks 1/W of all four passes), the fragments travel to the owner of their
input-row band in one all-to-all over NVLink, and the owner sorts and merges its band.  Nothing
is walked twice, and the result is still bit-identical to the single-GPU build.
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


def band_bounds(ncx: int, ncy: int, world_size: int) -> list[int]:
    """Flat input-cell bounds of the W input-row bands: ``[lo_0, lo_1, ..., lo_{W-1}, n_in]``."""
    return [band_cells(ncx, ncy, r, world_size)[0] for r in range(world_size)] + [ncx * ncy]


def _merge_band(rank: int, bounds: list[int], counts_by_src: torch.Tensor, recv_key, recv_val, n_in, n_out):
    return _device.build2d_merge(counts_by_src, recv_key, recv_val, bounds[rank], n_in, n_out)


def build_weights_2d_sharded(x_in, y_in, x_out, y_out, weights_input=None, replicate: bool = False,
                             group=None, device=None) -> _device.DeviceWeights:
    """Strong-scaling build of ONE large grid pair over the ranks of ``group``.

    Every rank holds the full coordinate arrays (134 MB at 2049^2) and walks every W-th block of 32 sweep
    lines of all four passes; its fragments come out bucketed by input cell, so the share of every
    input-row band is one contiguous range.  Two all-to-alls follow (the per-cell fragment counts, fixed
    size; then the fragments themselves, ~1 GB / W^2 per pair of ranks at 2048^2), and each rank merges its
    band (``rg_build2d_merge``).  Returns this rank's band of the public triplets; with ``replicate`` the
    bands are all-gathered so that every rank holds the full matrix (equal to the single-GPU build bit
    for bit)."""
    rank, W = world(group)
    if W == 1:
        return _device.build_weights_2d(x_in, y_in, x_out, y_out, weights_input, device=device)
    nxi, nyi = x_in.shape
    bounds = band_bounds(nxi - 1, nyi - 1, W)
    part = _device.build2d_part_walk(x_in, y_in, x_out, y_out, weights_input, rank, W, bounds, device=device)
    dev = part.frag_key.device
    band_sizes = [bounds[r + 1] - bounds[r] for r in range(W)]
    cb = band_sizes[rank]
    # 1. counts: my counts of band d -> rank d; I receive [W, cb]
    counts_by_src = torch.empty(W * cb, dtype=torch.int32, device=dev)
    dist.all_to_all_single(counts_by_src, part.counts, output_split_sizes=[cb] * W,
                           input_split_sizes=band_sizes, group=group)
    counts_by_src = counts_by_src.view(W, cb)
    recv_sizes = counts_by_src.sum(dim=1, dtype=torch.int64).cpu().tolist()  # the one host sync of the exchange
    send_sizes = [part.band_offsets[r + 1] - part.band_offsets[r] for r in range(W)]
    # 2. fragments
    n_recv = int(sum(recv_sizes))
    recv_key = torch.empty(n_recv, dtype=torch.int64, device=dev)
    recv_val = torch.empty(n_recv, dtype=torch.float64, device=dev)
    dist.all_to_all_single(recv_key, part.frag_key, output_split_sizes=recv_sizes, input_split_sizes=send_sizes,
                           group=group)
    dist.all_to_all_single(recv_val, part.frag_val, output_split_sizes=recv_sizes, input_split_sizes=send_sizes,
                           group=group)
    dw = _merge_band(rank, bounds, counts_by_src, recv_key, recv_val, part.n_in, part.n_out)
    dw.stats.update(part.check())
    if not replicate:
        return dw
    ii, io, v = allgather_concat([dw.indices_input, dw.indices_output, dw.values], group)
    out = _device.DeviceWeights(ii, io, v, dw.n_in, dw.n_out)
    out.stats = dw.stats
    return out


def build_weights_2d_sharded_local(x_in, y_in, x_out, y_out, weights_input=None, world_size: int = 2,
                                   device=None) -> list[_device.DeviceWeights]:
    """The line-sharded build with all ``world_size`` ranks played one after the other on ONE GPU and the
    all-to-all replaced by slicing: the same kernels and the same merge as ``build_weights_2d_sharded``,
    used to validate the partition on a single device (tests) and to time the per-rank share."""
    W = int(world_size)
    nxi, nyi = x_in.shape
    bounds = band_bounds(nxi - 1, nyi - 1, W)
    parts = [_device.build2d_part_walk(x_in, y_in, x_out, y_out, weights_input, r, W, bounds, device=device)
             for r in range(W)]
    out = []
    for d in range(W):
        lo, hi = bounds[d], bounds[d + 1]
        counts_by_src = torch.stack([p.counts[lo:hi] for p in parts])
        recv_key = torch.cat([p.frag_key[p.band_offsets[d]:p.band_offsets[d + 1]] for p in parts])
        recv_val = torch.cat([p.frag_val[p.band_offsets[d]:p.band_offsets[d + 1]] for p in parts])
        dw = _merge_band(d, bounds, counts_by_src, recv_key, recv_val, parts[0].n_in, parts[0].n_out)
        out.append(dw)
    for d, p in enumerate(parts):
        out[d].stats.update(p.check())
    return out
