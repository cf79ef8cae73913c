"""
Multi-GPU sharding (one process per GPU, ``torch.distributed`` over NCCL/NVLink).

The path shards without any data-path collective:
  * independent orthogonal slices (spectra of config 2, frames of configs 3/4, the per-slice grids of config 4:
    ``build_weights_2d_slices``) are split into contiguous rank ranges (``shard_range``);
  * ONE large 2D build (``build_weights_2d_sharded``) is split into contiguous INPUT-ROW bands: the public layout is
    sorted by input cell first (regridding/_weights/_weights_arrays.py:54-59), so the per-rank results concatenate
    in rank order with no re-sort.  ``exchange="band"`` (default): a rank walks only the sweep segments that can
    reach its band (``rg_build2d_band``) -- no fragment is exchanged and there is no collective at all: every walk
    state a rank uses is verified by the rank itself.  ``"p2p"`` / ``"nccl"``: the older line-sharded build (sweep lines dealt out across the ranks,
    fragments read by the band owners over NVLink peer memory or sent by all-to-all).  ``build_weights_2d_banded``:
    every rank walks everything and keeps its band.
All of them are bit-identical to the single-GPU build.  The optional all-gather of the band triplets
(``replicate=True``) is for callers who want the full matrix on every rank (e.g. before a frame-sharded apply).
"""

from __future__ import annotations

import numpy as _np
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


_ALLGATHER_CHOICE: dict = {}


def allgather_concat(tensors, group=None):
    """Variable-length all-gather along dim 0 of one tensor or of a list of 1-D 8-byte tensors that share their
    length: every rank gets ``cat([t_0, ..., t_{W-1}])`` of each.  One small collective for the lengths, then either
    ONE all-gather of the tensors packed side by side and padded to the longest band + one compaction per tensor
    (also the gloo / CPU path), or NCCL broadcasts straight into views of the result; on CUDA the first call times
    both and later calls take the faster one."""
    single = isinstance(tensors, torch.Tensor)
    ts = [tensors] if single else list(tensors)
    rank, W = world(group)
    if W == 1:
        return tensors
    dev = ts[0].device
    if any(t.ndim != 1 or t.element_size() != 8 or t.shape[0] != ts[0].shape[0] for t in ts):
        raise ValueError("allgather_concat expects 1-D tensors of 8-byte elements with one common length")
    n = torch.tensor([ts[0].shape[0]], dtype=torch.int64, device=dev)
    counts = torch.empty(W, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts_h = counts.cpu().tolist()
    most = max(counts_h)
    if most == 0:
        return tensors
    k = len(ts)
    offs = [0]
    for c in counts_h:
        offs.append(offs[-1] + c)

    def by_broadcasts():
        # NCCL gathers uneven shares straight into views of the result (one grouped broadcast per rank and tensor):
        # no padding, no compaction pass
        outs = []
        for t in ts:
            full = torch.empty(offs[-1], dtype=t.dtype, device=dev)
            dist.all_gather([full[offs[r]:offs[r + 1]] for r in range(W)], t.contiguous(), group=group)
            outs.append(full)
        return outs

    def by_padded_allgather():
        # ONE all_gather_into_tensor (NCCL's own all-gather: rings / NVLS) of the tensors packed side by side and
        # padded to the longest share, then one compaction (torch.cat of W views) per tensor
        packed = torch.empty((k, most), dtype=torch.int64, device=dev)
        for q, t in enumerate(ts):
            packed[q, :t.shape[0]] = t.contiguous().view(torch.int64)
        flat = torch.empty(W * k * most, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(flat, packed.view(-1), group=group)
        gathered = flat.view(W, k, most)
        return [torch.cat([gathered[r, q, :counts_h[r]] for r in range(W)]).view(t.dtype) for q, t in enumerate(ts)]

    if dev.type == "cuda":
        # Which of the two is faster depends on the box (NVLS, the NCCL version, message sizes: measured 0.8 and 2.5 ms
        # for the broadcasts on two 8 x B200 boxes): the first call of a (group size, tensor count, size class) times
        # both on the device, the ranks agree on the maximum, later calls take the winner.
        key = (W, k, int(most).bit_length(), dev.index)
        mode = _ALLGATHER_CHOICE.get(key)
        if mode is None:
            times, outs = [], None
            for fn in (by_broadcasts, by_padded_allgather):
                fn()  # warm-up (communicator set-up, allocator)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                outs = fn()
                e1.record()
                torch.cuda.synchronize(dev)
                times.append(e0.elapsed_time(e1))
            t = torch.tensor(times, dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            tb, tp = t.tolist()
            _ALLGATHER_CHOICE[key] = "broadcasts" if tb <= tp else "padded"
            return outs[0] if single else outs
        outs = by_broadcasts() if mode == "broadcasts" else by_padded_allgather()
        return outs[0] if single else outs
    packed = torch.empty((k, most), dtype=torch.int64, device=dev)
    for q, t in enumerate(ts):
        packed[q, :t.shape[0]] = t.contiguous().view(torch.int64)
    flat = torch.empty(W * k * most, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(flat, packed.view(-1), group=group)
    gathered = flat.view(W, k, most)
    outs = [torch.cat([gathered[r, q, :counts_h[r]] for r in range(W)]).view(t.dtype) for q, t in enumerate(ts)]
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


def build_weights_2d_slices(slices_host, weights_input=None, group=None, device=None, chunk: int = 4) -> dict:
    """Per-slice grids (BASELINE config 4; the reference's Python loop, _weights_conservative.py:110-139) sharded over
    the ranks with NO collective: rank r builds the contiguous share ``shard_range(n_slices, r, W)``.

    ``slices_host`` is a sequence of ``(x_in, y_in, x_out, y_out)`` host arrays (output coordinates already perturbed
    -- the seeded jitter is one serial NumPy stream by the reference's definition: draw it once, e.g. on rank 0, and
    hand every rank its slices).  The rank's slices go through pinned staging buffers; the H2D copy of chunk k + 1
    overlaps the builds of chunk k (``rg_build2d_batched``: no host synchronisation inside a chunk).
    Returns ``{slice index: DeviceWeights}`` of this rank's slices."""
    rank, W = world(group)
    dev = _device.cuda_device(device)
    n = len(slices_host)
    lo, hi = shard_range(n, rank, W)
    mine = list(range(lo, hi))
    out: dict = {}
    if not mine:
        return out
    copy_stream = torch.cuda.Stream(dev)

    def upload(ks):
        with torch.cuda.stream(copy_stream):
            t = [[torch.from_numpy(_np.ascontiguousarray(a, dtype=_np.float64)).pin_memory().to(dev, non_blocking=True)
                  for a in slices_host[k]] for k in ks]
            ev = torch.cuda.Event()
            ev.record()
        return t, ev

    chunks = [mine[c:c + chunk] for c in range(0, len(mine), chunk)]
    nxt = upload(chunks[0])
    for c, ks in enumerate(chunks):
        t, ev = nxt
        torch.cuda.current_stream(dev).wait_event(ev)
        if c + 1 < len(chunks):
            nxt = upload(chunks[c + 1])
        ws_in = None if weights_input is None else [weights_input[k] for k in ks]
        for k, dw in zip(ks, _device.build_weights_2d_batched(t, ws_in, device=dev)):
            out[k] = dw
    return out


def band_bounds(ncx: int, ncy: int, world_size: int) -> list[int]:
    """Flat input-cell bounds of the W input-row bands: ``[lo_0, lo_1, ..., lo_{W-1}, n_in]``."""
    return [band_cells(ncx, ncy, r, world_size)[0] for r in range(world_size)] + [ncx * ncy]


FRAG_BYTES = 16


class PeerArena:
    """Peer-mapped (symmetric) memory of one rank for the line-sharded build: the build workspace (it holds the
    per-cell fragment counts), the fragment buffer and a small header (band offsets), allocated with
    ``torch.distributed._symmetric_memory`` so that every rank of the group can read them in place over
    NVLink.  Rendezvous is collective and slow (milliseconds): arenas are cached per (group, shapes)."""

    def __init__(self, group, device: torch.device, shape: tuple[int, int, int, int], capacity: int, n_hdr: int):
        import torch.distributed._symmetric_memory as symm

        self.shape = shape
        self.capacity = int(capacity)
        self.ws_bytes = _device.build2d_workspace_bytes(*shape)
        al = lambda x: (x + 255) // 256 * 256  # noqa: E731
        self.off_frags = al(self.ws_bytes)
        self.off_hdr = self.off_frags + al(FRAG_BYTES * self.capacity)
        total = self.off_hdr + al(8 * n_hdr)
        self.n_hdr = n_hdr
        self.buf = symm.empty(total, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.base = [int(p) for p in self.hdl.buffer_ptrs]
        self.workspace = self.buf[:self.ws_bytes]
        self.frags = self.buf[self.off_frags:self.off_frags + FRAG_BYTES * self.capacity].view(torch.int64).view(-1, 2)
        self.header = self.buf[self.off_hdr:self.off_hdr + 8 * n_hdr].view(torch.int64)

    def peer_header(self, r: int) -> torch.Tensor:
        return self.hdl.get_buffer(r, (self.n_hdr,), torch.int64, self.off_hdr // 8)

    def barrier(self, channel: int = 0):
        self.hdl.barrier(channel=channel)


_arenas: dict = {}


def _arena(group, device, shape, W, min_capacity: int = 0) -> PeerArena:
    key = (id(group) if group is not None else 0, device.index, shape)
    a = _arenas.get(key)
    ci, co = (shape[0] - 1) * (shape[1] - 1), (shape[2] - 1) * (shape[3] - 1)
    # first guess: 16 fragments per cell of the finer grid, shared by W ranks, + 30 %; REGRID_B200_ARENA_FRAGS_PER_CELL
    # overrides the 16 (tests use a tiny value to exercise the growth path)
    import os

    per_cell = float(os.environ.get("REGRID_B200_ARENA_FRAGS_PER_CELL", "16"))
    want = max(int(min_capacity), int(1.3 * per_cell * max(ci, co) / W) + 64)
    if a is None or a.capacity < min_capacity:
        _arenas.pop(key, None)
        a = None  # release the old mapping before the new rendezvous
        a = PeerArena(group, device, shape, want, W + 1)
        _arenas[key] = a
    return a


def _sharded_build_p2p(x_in, y_in, x_out, y_out, weights_input, rank, W, bounds, group, device, mark):
    """Fused exchange: nothing is sent.  Every rank walks into its peer-mapped arena; after one barrier the band
    owners read the other ranks' counts and fragments in place (NVLink loads inside the gather kernels)."""
    dev = _device.cuda_device(device if device is not None else (x_in.device if isinstance(x_in, torch.Tensor) else None))
    shape = (int(x_in.shape[0]), int(x_in.shape[1]), int(x_out.shape[0]), int(x_out.shape[1]))
    need = 0
    while True:
        arena = _arena(group, dev, shape, W, need)
        part = _device.build2d_part_count(x_in, y_in, x_out, y_out, weights_input, rank, W, bounds, device=dev,
                                          workspace=arena.workspace, header=arena.header)
        if part.n_fragments <= arena.capacity:
            _device.build2d_part_fill(part, arena.frags)
        mark("walk")
        arena.barrier(0)  # every rank's walk and header are complete (stream-ordered)
        hdr = torch.stack([arena.peer_header(r) for r in range(W)]).cpu()  # [W, W + 1]; the one host sync
        most = int(hdr[:, W].max())
        if most <= arena.capacity:
            break
        arena.barrier(1)
        need = int(most * 1.25) + 4096  # same decision on every rank: grow the arenas together and walk again
    cb = bounds[rank + 1] - bounds[rank]
    sizes = [int(hdr[s, rank + 1] - hdr[s, rank]) for s in range(W)]
    chunk_ptrs = [arena.base[s] + arena.off_frags + FRAG_BYTES * int(hdr[s, rank]) for s in range(W)]
    count_ptrs = [arena.base[s] + part.counts_offset + 4 * bounds[rank] for s in range(W)]
    counts_by_src = _device.build2d_gather_counts(count_ptrs, cb, dev)
    mark("exchange")
    dw = _device.build2d_merge(counts_by_src, chunk_ptrs, sizes, bounds[rank], part.n_in, part.n_out)
    arena.barrier(1)  # nobody starts overwriting its arena while a peer still reads it
    dw.stats.update(part.check())
    mark("merge")
    return dw


def _sharded_build_nccl(x_in, y_in, x_out, y_out, weights_input, rank, W, bounds, group, device, mark):
    """Exchange by NCCL all-to-all: the per-cell counts (fixed sizes), then the 16-byte fragment records."""
    part = _device.build2d_part_walk(x_in, y_in, x_out, y_out, weights_input, rank, W, bounds, device=device)
    mark("walk")
    dev = part.frags.device
    band_sizes = [bounds[r + 1] - bounds[r] for r in range(W)]
    cb = band_sizes[rank]
    counts_by_src = torch.empty(W * cb, dtype=torch.int32, device=dev)
    dist.all_to_all_single(counts_by_src, part.counts, output_split_sizes=[cb] * W,
                           input_split_sizes=band_sizes, group=group)
    counts_by_src = counts_by_src.view(W, cb)
    recv_sizes = counts_by_src.sum(dim=1, dtype=torch.int64).cpu().tolist()  # the one host sync of the exchange
    send_sizes = [part.band_offsets[r + 1] - part.band_offsets[r] for r in range(W)]
    recv = _device.frags_empty(int(sum(recv_sizes)), dev)
    dist.all_to_all_single(recv, part.frags, output_split_sizes=recv_sizes, input_split_sizes=send_sizes, group=group)
    mark("exchange")
    offs = [0]
    for n in recv_sizes:
        offs.append(offs[-1] + n)
    ptrs = [recv.data_ptr() + FRAG_BYTES * offs[s] for s in range(W)]
    dw = _device.build2d_merge(counts_by_src, ptrs, recv_sizes, bounds[rank], part.n_in, part.n_out)
    dw.stats.update(part.check())
    mark("merge")
    return dw


def _sharded_build_band(x_in, y_in, x_out, y_out, weights_input, rank, W, group, device, mark):
    """Exchange-free band build: rank r builds the band of input rows ``shard_range(ncx, r, W)`` by walking only the
    sweep segments that can reach it (``rg_build2d_band``).  NO collective: every walk state a rank uses is verified by
    the rank itself (chain of walked segments + exact check of every run's first state), so a band that does not
    verify, or whose buffers were too small, is rebuilt by its own rank without involving the others."""
    nxi, nyi = x_in.shape
    lo, hi = shard_range(nxi - 1, rank, W)
    dev = _device.cuda_device(device if device is not None else (x_in.device if isinstance(x_in, torch.Tensor) else None))
    if hi <= lo:
        mark("band")
        e = torch.empty(0, dtype=torch.int64, device=dev)
        return _device.DeviceWeights(e, e.clone(), torch.empty(0, dtype=torch.float64, device=dev),
                                     (nxi - 1) * (nyi - 1), (x_out.shape[0] - 1) * (x_out.shape[1] - 1))
    status, dw = "mismatch", None
    for _ in range(4):
        # (finish(): the one host synchronisation of the build)
        dw, status = _device.build2d_band_enqueue(x_in, y_in, x_out, y_out, weights_input, lo, hi, device=dev).finish()
        if status != "capacity":
            break
    mark("band")
    if status != "ok":
        ncy = nyi - 1
        dw = _device.build_weights_2d(x_in, y_in, x_out, y_out, weights_input, cell_band=(lo * ncy, hi * ncy), device=dev)
        mark("fallback")
    return dw


def build_weights_2d_sharded(x_in, y_in, x_out, y_out, weights_input=None, replicate: bool = False,
                             group=None, device=None, exchange: str = "band",
                             phases: dict | None = None) -> _device.DeviceWeights:
    """Strong-scaling build of ONE large grid pair over the ranks of ``group``.

    Every rank holds the full coordinate arrays (134 MB at 2049^2).  ``exchange="band"`` (default): rank r builds its
    band of input rows and walks only the sweep segments that can reach it -- no fragments are exchanged at all
    (``rg_build2d_band``).  ``exchange="p2p"`` / ``"nccl"``: the older line-sharded build (every rank walks every W-th
    block of 32 sweep lines, the band owners gather the fragments over NVLink peer memory or by all-to-all and merge).
    Returns this rank's band of the public triplets; with ``replicate`` the bands are all-gathered so that
    every rank holds the full matrix (equal to the single-GPU build bit for bit).
    ``phases`` (development) receives the milliseconds of the phases."""
    rank, W = world(group)
    if W == 1:
        return _device.build_weights_2d(x_in, y_in, x_out, y_out, weights_input, device=device)
    if exchange not in ("band", "p2p", "nccl"):
        raise ValueError(f"exchange must be 'band', 'p2p' or 'nccl', got {exchange!r}")
    nxi, nyi = x_in.shape
    bounds = band_bounds(nxi - 1, nyi - 1, W)
    marks = []

    def mark(name):
        if phases is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    mark("start")
    if exchange == "band":
        dw = _sharded_build_band(x_in, y_in, x_out, y_out, weights_input, rank, W, group, device, mark)
    else:
        impl = _sharded_build_p2p if exchange == "p2p" else _sharded_build_nccl
        dw = impl(x_in, y_in, x_out, y_out, weights_input, rank, W, bounds, group, device, mark)
    if replicate:
        ii, io, v = allgather_concat([dw.indices_input, dw.indices_output, dw.values], group)
        out = _device.DeviceWeights(ii, io, v, dw.n_in, dw.n_out)
        out.stats = dw.stats
        dw = out
        mark("allgather")
    if phases is not None:
        torch.cuda.synchronize(dw.device)
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            phases[name] = phases.get(name, 0.0) + a.elapsed_time(b)
    return dw


def build_weights_2d_sharded_local(x_in, y_in, x_out, y_out, weights_input=None, world_size: int = 2,
                                   device=None) -> list[_device.DeviceWeights]:
    """The line-sharded build with all ``world_size`` ranks played one after the other on ONE GPU and the
    exchange replaced by pointers into the other "ranks'" buffers (what the p2p exchange does over NVLink):
    the same kernels and the same merge as ``build_weights_2d_sharded``, used to validate the partition on a
    single device (tests) and to time the per-rank share."""
    W = int(world_size)
    nxi, nyi = x_in.shape
    bounds = band_bounds(nxi - 1, nyi - 1, W)
    parts = [_device.build2d_part_walk(x_in, y_in, x_out, y_out, weights_input, r, W, bounds, device=device)
             for r in range(W)]
    dev = parts[0].workspace.device
    out = []
    for d in range(W):
        cb = bounds[d + 1] - bounds[d]
        counts = _device.build2d_gather_counts([p.counts.data_ptr() + 4 * bounds[d] for p in parts], cb, dev)
        sizes = [p.band_offsets[d + 1] - p.band_offsets[d] for p in parts]
        ptrs = [p.frags.data_ptr() + FRAG_BYTES * p.band_offsets[d] for p in parts]
        out.append(_device.build2d_merge(counts, ptrs, sizes, bounds[d], parts[0].n_in, parts[0].n_out))
    for d, p in enumerate(parts):
        out[d].stats.update(p.check())
    return out
