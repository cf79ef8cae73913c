"""
Device-level operators: thin, typed wrappers of the C ABI working on
``torch`` CUDA tensors.  Everything here runs on the GPU; there is no CPU path.

The public, reference-compatible API (``regridding_b200.weights`` etc.) is built on
these; they are also what ``bench.py`` times for the HBM-resident numbers.
"""

from __future__ import annotations

import ctypes
import dataclasses
import os
import threading
import warnings

import numpy as np
import torch

from . import _lib

F64 = torch.float64
I64 = torch.int64
I32 = torch.int32

# kernel launches of ours per operator call (checked against the ncu launch list in profiles/)
LAUNCHES_BUILD2D = 20  # area 1, bbox+boundaries 4, guess 2, starts / count / check / repair 4 (four passes per launch), scans 6, emit walk 1, sort 1, merge 1
# line-sharded build per rank: walk share 14 (area 1, boundaries 4, guess, starts, count, check, repair, scan 3, emit) + merge 13 (counts 1, transpose 1, scans 9, gather-sort 1, emit 1)
LAUNCHES_BUILD2D_SHARDED = 27
LAUNCHES_CSR = 6       # hist, scan 3, fill, rank
LAUNCHES_APPLY = 1     # one k_apply launch per call (up to 65535 frame tiles)


def cuda_device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.RegridB200Error("regridding_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.RegridB200Error(f"regridding_b200 runs on CUDA devices only, got {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def to_device(a, device: torch.device, dtype=F64) -> torch.Tensor:
    """numpy array or tensor -> contiguous tensor on `device`."""
    if isinstance(a, torch.Tensor):
        if a.device == device and a.dtype == dtype and a.is_contiguous():
            return a
        return a.to(device=device, dtype=dtype).contiguous()
    a = np.ascontiguousarray(a, dtype={F64: np.float64, I64: np.int64, I32: np.int32}[dtype])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)  # read-only (broadcast) inputs are only read
        return torch.from_numpy(a).to(device)


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def frags_empty(n: int, device) -> torch.Tensor:
    """Raw fragment records: 16 bytes each ({uint64 key, float64 weight}), viewed as int64 [n, 2]."""
    return torch.empty((max(int(n), 1), 2), dtype=I64, device=device)[:int(n)]


# ---------------------------------------------------------------------------
# weights container
# ---------------------------------------------------------------------------


@dataclasses.dataclass
class CSR:
    row_ptr: torch.Tensor  # int32 [n_out + 1]
    col: torch.Tensor      # int32 [nnz]
    val: torch.Tensor      # float64 [nnz]
    n_in: int
    n_out: int


@dataclasses.dataclass
class ApplyPlan:
    """Per-tile footprints + tile-local indices of one CSR between 2D cell grids (rg_apply_plan_build)."""

    csr: CSR
    shape_in: tuple[int, int]
    shape_out: tuple[int, int]
    tile_info: torch.Tensor
    tile_rows: torch.Tensor
    slot_val: torch.Tensor   # float64 [n_slots]: weights in the order the staged kernel consumes them
    slot_lidx: torch.Tensor  # int16 (uint16 bits) [n_slots]: byte offset of the referenced cell in a staged frame
    n_tiles: int
    n_generic_tiles: int


class DeviceWeights:
    """One set of weights resident in HBM: the public COO (sorted by (input, output),
    unique pairs, int64/int64/float64) plus, lazily, the CSR-by-output form the apply uses."""

    def __init__(self, ii: torch.Tensor, io: torch.Tensor, v: torch.Tensor, n_in: int, n_out: int):
        self.indices_input = ii
        self.indices_output = io
        self.values = v
        self.n_in = int(n_in)
        self.n_out = int(n_out)
        self._csr: CSR | None = None
        self._plans: dict = {}
        self.stats: dict | None = None

    @property
    def nnz(self) -> int:
        return int(self.values.numel())

    @property
    def device(self) -> torch.device:
        return self.values.device

    def device_bytes(self) -> int:
        """Device memory this object keeps alive (COO + CSR + apply plans)."""
        n = sum(t.numel() * t.element_size() for t in (self.indices_input, self.indices_output, self.values))
        if self._csr is not None:
            n += sum(t.numel() * t.element_size() for t in (self._csr.row_ptr, self._csr.col, self._csr.val))
        for p in self._plans.values():
            n += sum(t.numel() * t.element_size() for t in (p.tile_info, p.tile_rows, p.slot_val, p.slot_lidx))
        return n

    def csr(self) -> CSR:
        if self._csr is None:
            ii, io = self.indices_input, self.indices_output
            if self.nnz and (int(ii.min().item()) < 0 or int(io.min().item()) < 0):
                # negative (wrap-around) indices of descending 1D grids: Numba wraps them (rfw.py:179-182)
                ii = torch.where(ii < 0, ii + self.n_in, ii)
                io = torch.where(io < 0, io + self.n_out, io)
            self._csr = csr_from_coo(ii, io, self.values, self.n_in, self.n_out)
        return self._csr

    def plan(self, shape_in: tuple[int, int], shape_out: tuple[int, int]) -> ApplyPlan:
        key = (tuple(shape_in), tuple(shape_out))
        if key not in self._plans:
            self._plans[key] = build_apply_plan(self.csr(), key[0], key[1])
        return self._plans[key]

    def to_host(self) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """The reference's saved-weights element: ``(indices_input, indices_output, values)``."""
        return (self.indices_input.cpu().numpy(), self.indices_output.cpu().numpy(), self.values.cpu().numpy())

    @classmethod
    def from_host(cls, indices_input, indices_output, values, n_in: int, n_out: int, device=None) -> "DeviceWeights":
        device = cuda_device(device)
        ii = np.asarray(indices_input).astype(np.int64, copy=True)
        io = np.asarray(indices_output).astype(np.int64, copy=True)
        # negative (wrap-around) indices from descending 1D grids: numba wraps them (rfw.py:179-182)
        if ii.size and ii.min() < 0:
            ii[ii < 0] += n_in
        if io.size and io.min() < 0:
            io[io < 0] += n_out
        if ii.size and (ii.min() < 0 or ii.max() >= n_in or io.min() < 0 or io.max() >= n_out):
            raise IndexError("weights index out of range for the given shapes")
        v = np.ascontiguousarray(np.asarray(values), dtype=np.float64)
        return cls(to_device(ii, device, I64), to_device(io, device, I64), to_device(v, device, F64), n_in, n_out)


class HostWeights:
    """One saved-weights element that already lives on the host (``(indices_input, indices_output, values)``):
    what multi-slice builds return, so that device memory is bounded by one chunk of slices, not by their number."""

    def __init__(self, ii: np.ndarray, io: np.ndarray, v: np.ndarray, n_in: int, n_out: int):
        self.indices_input, self.indices_output, self.values = ii, io, v
        self.n_in, self.n_out = int(n_in), int(n_out)

    @property
    def nnz(self) -> int:
        return int(self.values.size)

    def to_host(self):
        return (self.indices_input, self.indices_output, self.values)


# ---------------------------------------------------------------------------
# 2D conservative build
# ---------------------------------------------------------------------------


def build_weights_2d(x_in, y_in, x_out, y_out, weights_input=None, cell_band: tuple[int, int] | None = None,
                     device=None) -> DeviceWeights:
    """weights_conservative_2d + _coalesce on the GPU (see ``rg_build2d_*`` in include/regrid_b200.h).

    All four coordinate arrays are 2D vertex grids; the OUTPUT coordinates must already
    carry the reference's host-side perturbation if one is wanted.
    """
    L = _lib.load()
    device = cuda_device(device if device is not None else (x_in.device if isinstance(x_in, torch.Tensor) else None))
    xi, yi = to_device(x_in, device), to_device(y_in, device)
    xo, yo = to_device(x_out, device), to_device(y_out, device)
    if xi.ndim != 2 or xi.shape != yi.shape or xo.ndim != 2 or xo.shape != yo.shape:
        raise ValueError("grids must be 2D arrays with matching x / y shapes")
    nxi, nyi = xi.shape
    nxo, nyo = xo.shape
    n_in, n_out = (nxi - 1) * (nyi - 1), (nxo - 1) * (nyo - 1)
    w = None
    if weights_input is not None:
        w = to_device(weights_input, device)
        if tuple(w.shape) != (nxi - 1, nyi - 1):
            raise ValueError(f"weights_input must have the input cell shape {(nxi - 1, nyi - 1)}, got {tuple(w.shape)}")
    lo, hi = (0, n_in) if cell_band is None else (int(cell_band[0]), int(cell_band[1]))

    # A REPEAT build of a shape (the longest bucket of an earlier build is known) walks every segment once: the whole grid
    # as one band through rg_build2d_band_onewalk (bit-identical; 3.2 ms against 3.9 ms at 2048^2).  Anything but "ok"
    # (buckets too small for these coordinates, a state that does not verify) falls through to the standard build.
    whole = (nxi, nyi, nxo, nyo, 0, nxi - 1)
    # (large grids only: below ~1024^2 cells the fixed cost of the band pipeline outweighs the saved walk)
    if (cell_band is None and n_in + n_out >= 2_000_000 and whole in _band_caps and _band_caps[whole][2] > 0
            and whole not in _band_onewalk_failed and not os.environ.get("REGRID_B200_BAND_TWO_WALKS")):
        dw, status = build2d_band_enqueue(xi, yi, xo, yo, w, 0, nxi - 1, device=device).finish()
        if status == "ok":
            return dw

    with torch.cuda.device(device):
        st = _stream(device)
        nbytes = ctypes.c_size_t()
        _lib.check(L.rg_build2d_workspace_bytes(nxi, nyi, nxo, nyo, ctypes.byref(nbytes)), "rg_build2d_workspace_bytes")
        ws = _workspace(nbytes.value, device)
        nfrag = ctypes.c_int64()
        _lib.check(L.rg_build2d_count(device.index, st, nxi, nyi, nxo, nyo, xi.data_ptr(), yi.data_ptr(),
                                      xo.data_ptr(), yo.data_ptr(), lo, hi, ws.data_ptr(), ws.numel(),
                                      ctypes.byref(nfrag)), "rg_build2d_count")
        nf = nfrag.value
        frags = frags_empty(nf, device)
        nnz_c = ctypes.c_int64()
        _lib.check(L.rg_build2d_fill(device.index, st, nxi, nyi, nxo, nyo, xi.data_ptr(), yi.data_ptr(),
                                     xo.data_ptr(), yo.data_ptr(), _lib.ptr(w), lo, hi, ws.data_ptr(), ws.numel(),
                                     frags.data_ptr(), nf, ctypes.byref(nnz_c)), "rg_build2d_fill")
        nnz = nnz_c.value
        ii = torch.empty(nnz, dtype=I64, device=device)
        io = torch.empty(nnz, dtype=I64, device=device)
        v = torch.empty(nnz, dtype=F64, device=device)
        _lib.check(L.rg_build2d_emit(device.index, st, nxi, nyi, nxo, nyo, lo, hi, ws.data_ptr(), ws.numel(),
                                     frags.data_ptr(), nf,
                                     ii.data_ptr(), io.data_ptr(), v.data_ptr(), nnz), "rg_build2d_emit")
        stats = (ctypes.c_int32 * 8)()
        _lib.check(L.rg_build2d_stats(device.index, st, nxi, nyi, nxo, nyo, ws.data_ptr(), stats), "rg_build2d_stats")
    dw = DeviceWeights(ii, io, v, n_in, n_out)
    dw.stats = {"fragments": nf, "nnz": nnz, "repaired_segments": int(stats[1]), "unknown_guesses": int(stats[2])}
    if cell_band is None and whole not in _band_caps and int(stats[6]) > 0:
        # the longest bucket (stats[6]) sizes the fixed-capacity buckets of the one-walk rebuilds of this shape
        longest = int(stats[6])
        _band_caps[whole] = (int(nf * 1.02) + 1024, int(nnz * 1.02) + 1024, (longest + max(4, longest // 4) + 1) // 2 * 2)
    return dw


# learned buffer sizes of band builds: (shape, band) -> (fragments, triplets) of the previous build
_band_caps: dict = {}
_band_buckets: dict = {}          # (device, stream, band cells, bucket capacity) -> strided fragment buckets
_band_onewalk_failed: set = set()   # shapes whose one-walk build overflowed its buckets: two walks from then on
# scratch of the last band build shape: (workspace, fragment buffer); stream-ordered reuse on the current stream
_band_scratch: dict = {}


_pinned = threading.local()


def _pinned_counts() -> torch.Tensor:
    """pinned int64[8] the status counters of a band build are copied into (one per host thread: the copy is consumed
    before the thread finishes its next build)"""
    t = getattr(_pinned, "counts", None)
    if t is None:
        t = torch.empty(8, dtype=I64).pin_memory()
        _pinned.counts = t
    return t


def release_build_scratch() -> None:
    """Drop the device buffers the band builds keep between calls (workspace, fragment buffer, the fixed-capacity buckets
    of the one-walk builds: ~4 GB after a repeat build at 2048^2).  The learned sizes stay; the next build reallocates."""
    _band_scratch.clear()
    _band_buckets.clear()


@dataclasses.dataclass
class BandBuild:
    """One enqueued ``rg_build2d_band``: nothing has been synchronised yet.  ``counts`` (device int64[8]) holds
    fragments, triplets and the status flags once the stream has run; ``finish()`` reads it."""

    ii: torch.Tensor
    io: torch.Tensor
    v: torch.Tensor
    counts: torch.Tensor
    frag_capacity: int
    nnz_capacity: int
    n_in: int
    n_out: int
    key: tuple
    keep: tuple  # buffers the enqueued kernels still use
    bucket_capacity: int = 0   # > 0: this is a one-walk build (rg_build2d_band_onewalk)

    MISMATCH, CAPACITY = 6, 7

    def finish(self, counts_host=None):
        """-> (DeviceWeights | None, status): status is "ok", "mismatch" (use the sequentially verified build) or
        "capacity" (buffers were too small: the learned sizes are updated, build again)."""
        if counts_host is None:
            pin = _pinned_counts()
            pin.copy_(self.counts, non_blocking=True)
            torch.cuda.current_stream(self.counts.device).synchronize()
            c = pin.tolist()
        else:
            c = counts_host.tolist()
        nfrag, nnz = int(c[0]), int(c[1])
        if c[2] or c[5]:
            raise _lib.RegridB200Error("rg_build2d_band: a sweep walk did not terminate (degenerate or folded grid)")
        if c[self.CAPACITY]:
            # (a one-walk build whose buckets overflowed is repeated as a two-walk build, and so are all later
            # builds of the shape)
            if self.bucket_capacity > 0:
                _band_onewalk_failed.add(self.key)
            _band_caps[self.key] = (max(int(nfrag * 1.05) + 1024, self.frag_capacity),
                                    max(int(nfrag * 0.55) + 1024, self.nnz_capacity), 0)
            return None, "capacity"
        if c[self.MISMATCH]:
            return None, "mismatch"
        # the longest bucket sizes the fixed-capacity buckets of the one-walk rebuilds of this shape (even: 32-byte rows)
        longest = int(c[3])
        bcap = 0 if longest <= 0 else (longest + max(4, longest // 4) + 1) // 2 * 2
        if self.key in _band_onewalk_failed:
            bcap = 0
        _band_caps[self.key] = (int(nfrag * 1.02) + 1024, int(nnz * 1.02) + 1024, bcap)
        ii, io, v = self.ii[:nnz], self.io[:nnz], self.v[:nnz]
        if self.nnz_capacity > 1.5 * nnz + 4096:   # first build of a shape: do not keep the over-sized estimate alive
            ii, io, v = ii.clone(), io.clone(), v.clone()
        dw = DeviceWeights(ii, io, v, self.n_in, self.n_out)
        dw.stats = {"fragments": nfrag, "nnz": nnz, "repaired_segments": 0, "unknown_guesses": int(c[4])}
        return dw, "ok"


def build2d_band_enqueue(x_in, y_in, x_out, y_out, weights_input, row_lo: int, row_hi: int, device=None) -> BandBuild:
    """Enqueue the band build of input rows ``[row_lo, row_hi)`` (``rg_build2d_band``); no host synchronisation."""
    L = _lib.load()
    device = cuda_device(device if device is not None else (x_in.device if isinstance(x_in, torch.Tensor) else None))
    xi, yi = to_device(x_in, device), to_device(y_in, device)
    xo, yo = to_device(x_out, device), to_device(y_out, device)
    if xi.ndim != 2 or xi.shape != yi.shape or xo.ndim != 2 or xo.shape != yo.shape:
        raise ValueError("grids must be 2D arrays with matching x / y shapes")
    nxi, nyi = xi.shape
    nxo, nyo = xo.shape
    n_in, n_out = (nxi - 1) * (nyi - 1), (nxo - 1) * (nyo - 1)
    w = None
    if weights_input is not None:
        w = to_device(weights_input, device)
        if tuple(w.shape) != (nxi - 1, nyi - 1):
            raise ValueError(f"weights_input must have the input cell shape {(nxi - 1, nyi - 1)}, got {tuple(w.shape)}")
    key = (nxi, nyi, nxo, nyo, int(row_lo), int(row_hi))
    nb = (int(row_hi) - int(row_lo)) * (nyi - 1)
    bcap = 0   # (0: two walks -- no build of this shape has reported its longest bucket yet, or a one-walk build overflowed)
    if key in _band_caps:
        fcap, ncap, bcap = _band_caps[key]
    else:
        share = nb / max(n_in, 1)
        fcap = int(10 * (nb + n_out * share)) + 4096
        ncap = fcap // 2
    if os.environ.get("REGRID_B200_BAND_TWO_WALKS") or nb * bcap * 16 > (8 << 30):
        bcap = 0
    with torch.cuda.device(device):
        st = _stream(device)
        # workspace and fragment buffer are scratch: kept per (device, shape) so that repeated builds allocate nothing
        skey = (device.index, st, nxi, nyi, nxo, nyo)
        scratch = _band_scratch.get(skey)
        if scratch is None or scratch[1].shape[0] < fcap:
            _band_scratch.clear()
            scratch = (_workspace(build2d_workspace_bytes(nxi, nyi, nxo, nyo), device), frags_empty(fcap, device))
            _band_scratch[skey] = scratch
        ws, frags = scratch
        # one allocation: counts[8] | indices_input | indices_output | values (views; 8-byte elements)
        # (every array starts on a 256-byte boundary, like separate allocations would)
        pad = (ncap + 31) // 32 * 32
        buf = torch.empty(32 + 3 * pad, dtype=I64, device=device)
        counts, ii, io = buf[:8], buf[32:32 + ncap], buf[32 + pad:32 + pad + ncap]
        v = buf[32 + 2 * pad:32 + 2 * pad + ncap].view(F64)
        strided = None
        if bcap > 0:
            # ONE walk into fixed-capacity buckets (capacity learned from an earlier build of this shape)
            bkey = (device.index, st, nb, bcap)
            strided = _band_buckets.get(bkey)
            if strided is None:
                _band_buckets.clear()
                strided = frags_empty(nb * bcap, device)
                _band_buckets[bkey] = strided
            _lib.check(L.rg_build2d_band_onewalk(device.index, st, nxi, nyi, nxo, nyo, xi.data_ptr(), yi.data_ptr(),
                                                 xo.data_ptr(), yo.data_ptr(), _lib.ptr(w), int(row_lo), int(row_hi),
                                                 ws.data_ptr(), ws.numel(), frags.data_ptr(), fcap,
                                                 ii.data_ptr(), io.data_ptr(), v.data_ptr(), ncap, counts.data_ptr(),
                                                 strided.data_ptr(), bcap), "rg_build2d_band_onewalk")
        else:
            _lib.check(L.rg_build2d_band(device.index, st, nxi, nyi, nxo, nyo, xi.data_ptr(), yi.data_ptr(),
                                         xo.data_ptr(), yo.data_ptr(), _lib.ptr(w), int(row_lo), int(row_hi),
                                         ws.data_ptr(), ws.numel(), frags.data_ptr(), fcap,
                                         ii.data_ptr(), io.data_ptr(), v.data_ptr(), ncap, counts.data_ptr()),
                       "rg_build2d_band")
    return BandBuild(ii, io, v, counts, fcap, ncap, n_in, n_out, key, (ws, frags, xi, yi, xo, yo, w, strided), bcap)


def build_weights_2d_band(x_in, y_in, x_out, y_out, weights_input=None, row_band: tuple[int, int] | None = None,
                          device=None) -> DeviceWeights:
    """The band ``row_band = (row_lo, row_hi)`` of input rows of the 2D conservative weights, by the exchange-free
    band build; falls back to the sequentially verified banded build if a chain of walk states does not verify."""
    nxi, nyi = x_in.shape
    lo, hi = (0, nxi - 1) if row_band is None else (int(row_band[0]), int(row_band[1]))
    if lo >= hi:
        dev = cuda_device(device if device is not None else (x_in.device if isinstance(x_in, torch.Tensor) else None))
        e = torch.empty(0, dtype=I64, device=dev)
        return DeviceWeights(e, e.clone(), torch.empty(0, dtype=F64, device=dev), (nxi - 1) * (nyi - 1),
                             (x_out.shape[0] - 1) * (x_out.shape[1] - 1))
    for _ in range(3):
        dw, status = build2d_band_enqueue(x_in, y_in, x_out, y_out, weights_input, lo, hi, device=device).finish()
        if status == "ok":
            return dw
        if status == "mismatch":
            break
    return build_weights_2d(x_in, y_in, x_out, y_out, weights_input, cell_band=(lo * (nyi - 1), hi * (nyi - 1)),
                            device=device)


def build_weights_2d_batched(slices, weights_input=None, device=None) -> list:
    """Per-slice builds (every slice its own grid pair, all of one shape): ``slices`` is a sequence of
    ``(x_in, y_in, x_out, y_out)`` on the device; the builds are enqueued back to back with no host synchronisation
    (``rg_build2d_batched``) and the counts are read once at the end.  Returns one ``DeviceWeights`` per slice; a
    slice whose buffers were too small or whose walk states did not verify is rebuilt by ``build_weights_2d``."""
    L = _lib.load()
    slices = list(slices)
    if not slices:
        return []
    device = cuda_device(device if device is not None else slices[0][0].device)
    S = len(slices)
    t = [[to_device(a, device) for a in sl] for sl in slices]
    nxi, nyi = t[0][0].shape
    nxo, nyo = t[0][2].shape
    for sl in t:
        if sl[0].shape != (nxi, nyi) or sl[1].shape != (nxi, nyi) or sl[2].shape != (nxo, nyo) or sl[3].shape != (nxo, nyo):
            raise ValueError("all slices of a batched build must share one input and one output grid shape")
    n_in, n_out = (nxi - 1) * (nyi - 1), (nxo - 1) * (nyo - 1)
    w = None
    if weights_input is not None:
        w = [to_device(wi, device) for wi in weights_input]
    key = (nxi, nyi, nxo, nyo, 0, nxi - 1)
    if key in _band_caps:
        fcap, ncap = _band_caps[key][:2]
    else:
        fcap = int(10 * (n_in + n_out)) + 4096
        ncap = fcap // 2
    arr = lambda ptrs: (ctypes.c_void_p * S)(*ptrs)  # noqa: E731
    with torch.cuda.device(device):
        st = _stream(device)
        ws = _workspace(build2d_workspace_bytes(nxi, nyi, nxo, nyo), device)
        frags = frags_empty(fcap, device)
        ii = [torch.empty(ncap, dtype=I64, device=device) for _ in range(S)]
        io = [torch.empty(ncap, dtype=I64, device=device) for _ in range(S)]
        v = [torch.empty(ncap, dtype=F64, device=device) for _ in range(S)]
        counts = torch.empty((S, 8), dtype=I64, device=device)
        _lib.check(L.rg_build2d_batched(device.index, st, S, nxi, nyi, nxo, nyo,
                                        arr([sl[0].data_ptr() for sl in t]), arr([sl[1].data_ptr() for sl in t]),
                                        arr([sl[2].data_ptr() for sl in t]), arr([sl[3].data_ptr() for sl in t]),
                                        None if w is None else arr([wi.data_ptr() for wi in w]),
                                        ws.data_ptr(), ws.numel(), frags.data_ptr(), fcap,
                                        arr([a.data_ptr() for a in ii]), arr([a.data_ptr() for a in io]),
                                        arr([a.data_ptr() for a in v]), ncap, counts.data_ptr()),
                   "rg_build2d_batched")
        host = counts.cpu()  # the one synchronisation
    out = []
    for s in range(S):
        bb = BandBuild(ii[s], io[s], v[s], counts[s], fcap, ncap, n_in, n_out, key, ())
        dw, status = bb.finish(host[s])
        if status != "ok":
            dw = build_weights_2d(*t[s], None if w is None else w[s], device=device)
        out.append(dw)
    return out


@dataclasses.dataclass
class PartFragments:
    """Fragments one rank produced by walking its share of the sweep lines (line-sharded build):
    bucketed by input cell, so every input-row band is the contiguous range
    ``[band_offsets[b], band_offsets[b + 1])`` of ``frags``."""

    counts: torch.Tensor        # int32 [n_in]: fragments per input cell (view into the workspace)
    counts_offset: int          # byte offset of ``counts`` in the workspace
    n_fragments: int
    band_offsets: list[int]     # host: first fragment of every band bound
    n_in: int
    n_out: int
    workspace: torch.Tensor
    shape: tuple[int, int, int, int]
    args: tuple                 # what rg_build2d_part_fill needs again
    frags: torch.Tensor | None = None   # int64 [n_fragments, 2] 16-byte records (after ``build2d_part_fill``)

    def check(self) -> dict:
        """Synchronises; raises if a walk of this rank did not terminate."""
        L = _lib.load()
        device = self.workspace.device
        stats = (ctypes.c_int32 * 8)()
        with torch.cuda.device(device):
            _lib.check(L.rg_build2d_stats(device.index, _stream(device), *self.shape, self.workspace.data_ptr(), stats),
                       "rg_build2d_stats")
        if stats[0] or stats[3]:
            raise _lib.RegridB200Error("rg_build2d_part_fill: a sweep walk did not terminate (degenerate or folded grid)")
        return {"fragments_walked": int(self.n_fragments), "repaired_segments": int(stats[1]),
                "unknown_guesses": int(stats[2])}


def build2d_workspace_bytes(nxi: int, nyi: int, nxo: int, nyo: int) -> int:
    nbytes = ctypes.c_size_t()
    _lib.check(_lib.load().rg_build2d_workspace_bytes(nxi, nyi, nxo, nyo, ctypes.byref(nbytes)),
               "rg_build2d_workspace_bytes")
    return int(nbytes.value)


def build2d_part_count(x_in, y_in, x_out, y_out, weights_input, part_rank: int, part_world: int,
                       cell_bounds: list[int], device=None, workspace: torch.Tensor | None = None,
                       header: torch.Tensor | None = None) -> PartFragments:
    """Rank ``part_rank`` of ``part_world``: locate, walk and count every ``part_world``-th block of 32 sweep
    lines of all four passes (``rg_build2d_part_count``).  ``workspace`` / ``header`` (int64, one entry per
    bound) may live in peer-mapped memory so that the other ranks can read the counts and offsets in place."""
    L = _lib.load()
    device = cuda_device(device if device is not None else (x_in.device if isinstance(x_in, torch.Tensor) else None))
    xi, yi = to_device(x_in, device), to_device(y_in, device)
    xo, yo = to_device(x_out, device), to_device(y_out, device)
    if xi.ndim != 2 or xi.shape != yi.shape or xo.ndim != 2 or xo.shape != yo.shape:
        raise ValueError("grids must be 2D arrays with matching x / y shapes")
    nxi, nyi = xi.shape
    nxo, nyo = xo.shape
    n_in, n_out = (nxi - 1) * (nyi - 1), (nxo - 1) * (nyo - 1)
    w = None
    if weights_input is not None:
        w = to_device(weights_input, device)
        if tuple(w.shape) != (nxi - 1, nyi - 1):
            raise ValueError(f"weights_input must have the input cell shape {(nxi - 1, nyi - 1)}, got {tuple(w.shape)}")
    nb = len(cell_bounds)
    bounds = (ctypes.c_int64 * nb)(*[int(b) for b in cell_bounds])
    offsets = (ctypes.c_int64 * nb)()
    if header is not None and header.numel() < nb:
        raise ValueError("header too small for the band bounds")
    with torch.cuda.device(device):
        st = _stream(device)
        ws = workspace if workspace is not None else _workspace(build2d_workspace_bytes(nxi, nyi, nxo, nyo), device)
        nfrag = ctypes.c_int64()
        coff = ctypes.c_size_t()
        _lib.check(L.rg_build2d_part_count(device.index, st, nxi, nyi, nxo, nyo, xi.data_ptr(), yi.data_ptr(),
                                           xo.data_ptr(), yo.data_ptr(), part_rank, part_world,
                                           ws.data_ptr(), ws.numel(), ctypes.byref(nfrag),
                                           nb, bounds, offsets, ctypes.byref(coff), _lib.ptr(header)),
                   "rg_build2d_part_count")
    counts = ws[coff.value:coff.value + 4 * n_in].view(I32)
    return PartFragments(counts, int(coff.value), int(nfrag.value), [int(o) for o in offsets], n_in, n_out, ws,
                         (nxi, nyi, nxo, nyo), (xi, yi, xo, yo, w, part_rank, part_world))


def build2d_part_fill(part: PartFragments, frags_buffer: torch.Tensor | None = None) -> PartFragments:
    """Emit walk of the rank's lines into ``frags_buffer`` (>= n_fragments records) or a fresh buffer."""
    L = _lib.load()
    xi, yi, xo, yo, w, part_rank, part_world = part.args
    device = part.workspace.device
    nf = part.n_fragments
    if frags_buffer is None:
        frags = frags_empty(nf, device)
    else:
        if frags_buffer.shape[0] < nf:
            raise ValueError("fragment buffer too small")
        frags = frags_buffer[:nf]
    nxi, nyi, nxo, nyo = part.shape
    with torch.cuda.device(device):
        _lib.check(L.rg_build2d_part_fill(device.index, _stream(device), nxi, nyi, nxo, nyo, xi.data_ptr(), yi.data_ptr(),
                                          xo.data_ptr(), yo.data_ptr(), _lib.ptr(w), part_rank, part_world,
                                          part.workspace.data_ptr(), part.workspace.numel(), frags.data_ptr(), nf),
                   "rg_build2d_part_fill")
    part.frags = frags
    return part


def build2d_part_walk(x_in, y_in, x_out, y_out, weights_input, part_rank: int, part_world: int,
                      cell_bounds: list[int], device=None) -> PartFragments:
    """``build2d_part_count`` + ``build2d_part_fill`` with library-allocated buffers."""
    return build2d_part_fill(build2d_part_count(x_in, y_in, x_out, y_out, weights_input, part_rank, part_world,
                                                cell_bounds, device=device))


def build2d_gather_counts(count_ptrs: list[int], n_cells: int, device) -> torch.Tensor:
    """``counts[s][c]`` from one device (or peer) pointer per source (``rg_build2d_gather_counts``)."""
    L = _lib.load()
    n_src = len(count_ptrs)
    out = torch.empty((n_src, max(n_cells, 1)), dtype=I32, device=device)[:, :n_cells]
    arr = (ctypes.c_void_p * n_src)(*count_ptrs)
    with torch.cuda.device(device):
        _lib.check(L.rg_build2d_gather_counts(device.index, _stream(device), n_cells, n_src, arr, out.data_ptr()),
                   "rg_build2d_gather_counts")
    return out


def build2d_merge(counts: torch.Tensor, chunk_ptrs: list[int], chunk_sizes: list[int], cell_offset: int,
                  n_in: int, n_out: int) -> DeviceWeights:
    """Band owner: ``counts`` int32 ``[n_src, n_band_cells]`` and one chunk of 16-byte fragment records per source
    (device or peer pointer + record count) -> the band's public triplets (``rg_build2d_merge`` /
    ``rg_build2d_merge_emit``)."""
    L = _lib.load()
    device = counts.device
    n_src, n_cells = int(counts.shape[0]), int(counts.shape[1])
    counts = counts.contiguous()
    n_recv = int(sum(chunk_sizes))
    ptrs = (ctypes.c_void_p * n_src)(*[int(p) if int(s) else None for p, s in zip(chunk_ptrs, chunk_sizes)])
    sizes = (ctypes.c_int64 * n_src)(*[int(s) for s in chunk_sizes])
    with torch.cuda.device(device):
        st = _stream(device)
        nbytes = ctypes.c_size_t()
        _lib.check(L.rg_build2d_merge_workspace_bytes(n_cells, n_src, ctypes.byref(nbytes)),
                   "rg_build2d_merge_workspace_bytes")
        ws = _workspace(nbytes.value, device)
        frags = frags_empty(n_recv, device)
        nnz_c = ctypes.c_int64()
        _lib.check(L.rg_build2d_merge(device.index, st, n_cells, n_src, counts.data_ptr(), ptrs, sizes,
                                      ws.data_ptr(), ws.numel(), frags.data_ptr(), ctypes.byref(nnz_c)),
                   "rg_build2d_merge")
        nnz = nnz_c.value
        ii = torch.empty(nnz, dtype=I64, device=device)
        io = torch.empty(nnz, dtype=I64, device=device)
        v = torch.empty(nnz, dtype=F64, device=device)
        _lib.check(L.rg_build2d_merge_emit(device.index, st, n_cells, n_src, int(cell_offset), ws.data_ptr(), ws.numel(),
                                           frags.data_ptr(), ii.data_ptr(), io.data_ptr(), v.data_ptr(), nnz),
                   "rg_build2d_merge_emit")
    dw = DeviceWeights(ii, io, v, n_in, n_out)
    dw.stats = {"fragments": n_recv, "nnz": nnz}
    return dw


def grid_area(x, y, device=None) -> torch.Tensor:
    L = _lib.load()
    device = cuda_device(device)
    xd, yd = to_device(x, device), to_device(y, device)
    nx, ny = xd.shape
    out = torch.empty((nx - 1, ny - 1), dtype=F64, device=device)
    with torch.cuda.device(device):
        _lib.check(L.rg_grid_area(device.index, _stream(device), nx, ny, xd.data_ptr(), yd.data_ptr(), out.data_ptr()),
                   "rg_grid_area")
    return out


def cell_length_1d(x: torch.Tensor) -> torch.Tensor:
    """(S, n) edge stacks -> (S, n-1) cell lengths x[i+1] - x[i]."""
    L = _lib.load()
    device = x.device
    S, n = x.shape
    out = torch.empty((S, n - 1), dtype=F64, device=device)
    with torch.cuda.device(device):
        _lib.check(L.rg_cell_length_1d(device.index, _stream(device), S, n, x.data_ptr(), out.data_ptr()),
                   "rg_cell_length_1d")
    return out


def transpose_conservative(dw: "DeviceWeights", volume_input: torch.Tensor, volume_output: torch.Tensor,
                           weights_input: torch.Tensor | None = None) -> torch.Tensor:
    """Values of the conservatively transposed weights (same order as ``dw``'s triplets)."""
    L = _lib.load()
    device = dw.device
    out = torch.empty(dw.nnz, dtype=F64, device=device)
    with torch.cuda.device(device):
        _lib.check(L.rg_transpose_conservative(device.index, _stream(device), dw.nnz, dw.n_in, dw.n_out,
                                               dw.indices_input.data_ptr(), dw.indices_output.data_ptr(),
                                               dw.values.data_ptr(), volume_input.data_ptr(), volume_output.data_ptr(),
                                               _lib.ptr(weights_input), out.data_ptr()),
                   "rg_transpose_conservative")
    return out


# ---------------------------------------------------------------------------
# apply
# ---------------------------------------------------------------------------


def csr_from_coo(ii: torch.Tensor, io: torch.Tensor, v: torch.Tensor, n_in: int, n_out: int) -> CSR:
    L = _lib.load()
    device = v.device
    nnz = int(v.numel())
    row_ptr = torch.empty(n_out + 1, dtype=I32, device=device)
    col = torch.empty(max(nnz, 1), dtype=I32, device=device)
    val = torch.empty(max(nnz, 1), dtype=F64, device=device)
    with torch.cuda.device(device):
        nbytes = ctypes.c_size_t()
        _lib.check(L.rg_csr_workspace_bytes(nnz, n_out, ctypes.byref(nbytes)), "rg_csr_workspace_bytes")
        ws = _workspace(nbytes.value, device)
        _lib.check(L.rg_csr_from_coo(device.index, _stream(device), nnz, n_in, n_out,
                                     ii.data_ptr(), io.data_ptr(), v.data_ptr(),
                                     row_ptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                     ws.data_ptr(), ws.numel()), "rg_csr_from_coo")
    return CSR(row_ptr, col[:nnz], val[:nnz], int(n_in), int(n_out))


def apply_csr(csr: CSR, values_in: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """values_in (F, n_in) float64 contiguous on the CSR's device -> (F, n_out)."""
    L = _lib.load()
    device = csr.val.device
    if values_in.dtype != F64 or not values_in.is_contiguous() or values_in.device != device:
        raise ValueError("values_in must be a contiguous float64 tensor on the weights' device")
    F, n_in = values_in.shape
    if n_in != csr.n_in:
        raise ValueError(f"values_in has {n_in} cells per frame, the weights expect {csr.n_in}")
    if out is None:
        out = torch.empty((F, csr.n_out), dtype=F64, device=device)
    elif out.shape != (F, csr.n_out) or out.dtype != F64 or not out.is_contiguous() or out.device != device:
        raise ValueError("out must be a contiguous float64 (F, n_out) tensor on the weights' device")
    with torch.cuda.device(device):
        _lib.check(L.rg_apply_csr(device.index, _stream(device), F, csr.n_in, csr.n_out,
                                  csr.row_ptr.data_ptr(), csr.col.data_ptr(), csr.val.data_ptr(),
                                  values_in.data_ptr(), out.data_ptr()), "rg_apply_csr")
    return out


def build_apply_plan(csr: CSR, shape_in: tuple[int, int], shape_out: tuple[int, int]) -> ApplyPlan:
    L = _lib.load()
    device = csr.val.device
    (h_in, w_in), (h_out, w_out) = (int(s) for s in shape_in), (int(s) for s in shape_out)
    if h_in * w_in != csr.n_in or h_out * w_out != csr.n_out:
        raise ValueError("plan shapes do not match the weights")
    nt, ni, nr = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    _lib.check(L.rg_apply_plan_sizes(h_out, w_out, ctypes.byref(nt), ctypes.byref(ni), ctypes.byref(nr)),
               "rg_apply_plan_sizes")
    tile_info = torch.empty(ni.value, dtype=I32, device=device)
    tile_rows = torch.empty(nr.value, dtype=I32, device=device)
    nnz = int(csr.val.numel())
    lidx = torch.empty(max(nnz, 1), dtype=torch.int16, device=device)  # scratch: tile-local index of every entry
    ng, ns = ctypes.c_int64(), ctypes.c_int64()
    with torch.cuda.device(device):
        _lib.check(L.rg_apply_plan_build(device.index, _stream(device), nnz, h_in, w_in, h_out, w_out,
                                         csr.row_ptr.data_ptr(), csr.col.data_ptr(), tile_info.data_ptr(),
                                         tile_rows.data_ptr(), lidx.data_ptr(), ctypes.byref(ng), ctypes.byref(ns)),
                   "rg_apply_plan_build")
        slot_val = torch.empty(max(ns.value, 8), dtype=F64, device=device)
        slot_lidx = torch.empty(max(ns.value, 8), dtype=torch.int16, device=device)
        _lib.check(L.rg_apply_plan_slots(device.index, _stream(device), h_out, w_out, csr.row_ptr.data_ptr(),
                                         csr.val.data_ptr(), tile_info.data_ptr(), lidx.data_ptr(), ns.value,
                                         slot_val.data_ptr(), slot_lidx.data_ptr()), "rg_apply_plan_slots")
        torch.cuda.current_stream(device).synchronize()  # lidx (scratch) is released on return
    return ApplyPlan(csr, (h_in, w_in), (h_out, w_out), tile_info, tile_rows, slot_val, slot_lidx, nt.value, ng.value)


def apply_planned(plan: ApplyPlan, values_in: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Shared-memory staged apply: values_in (F, n_in) -> (F, n_out); same bits as ``apply_csr``."""
    L = _lib.load()
    csr = plan.csr
    device = csr.val.device
    if values_in.dtype != F64 or not values_in.is_contiguous() or values_in.device != device:
        raise ValueError("values_in must be a contiguous float64 tensor on the weights' device")
    F, n_in = values_in.shape
    if n_in != csr.n_in:
        raise ValueError(f"values_in has {n_in} cells per frame, the weights expect {csr.n_in}")
    if out is None:
        out = torch.empty((F, csr.n_out), dtype=F64, device=device)
    elif out.shape != (F, csr.n_out) or out.dtype != F64 or not out.is_contiguous() or out.device != device:
        raise ValueError("out must be a contiguous float64 (F, n_out) tensor on the weights' device")
    with torch.cuda.device(device):
        _lib.check(L.rg_apply_planned(device.index, _stream(device), F, plan.shape_in[0], plan.shape_in[1],
                                      plan.shape_out[0], plan.shape_out[1], csr.row_ptr.data_ptr(),
                                      csr.col.data_ptr(), csr.val.data_ptr(), plan.tile_info.data_ptr(),
                                      plan.tile_rows.data_ptr(), plan.slot_val.data_ptr(), plan.slot_lidx.data_ptr(),
                                      plan.n_generic_tiles,
                                      values_in.data_ptr(), out.data_ptr()), "rg_apply_planned")
    return out


# ---------------------------------------------------------------------------
# 1D
# ---------------------------------------------------------------------------


def cons1d_batched(x_in: torch.Tensor, x_out: torch.Tensor, weights_input: torch.Tensor | None = None):
    """(S, n) / (S, m) edge stacks -> (ii, io, v) of shape (S, n + m) and counts (S,), emission order."""
    L = _lib.load()
    device = x_in.device
    S, n = x_in.shape
    m = x_out.shape[1]
    cap = n + m
    ii = torch.empty((S, cap), dtype=I64, device=device)
    io = torch.empty((S, cap), dtype=I64, device=device)
    v = torch.empty((S, cap), dtype=F64, device=device)
    counts = torch.empty(S, dtype=I64, device=device)
    with torch.cuda.device(device):
        _lib.check(L.rg_cons1d_batched(device.index, _stream(device), S, n, m, x_in.data_ptr(), x_out.data_ptr(),
                                       _lib.ptr(weights_input), ii.data_ptr(), io.data_ptr(), v.data_ptr(),
                                       counts.data_ptr()), "rg_cons1d_batched")
    return ii, io, v, counts


def regrid1d_conservative(x_in: torch.Tensor, x_out: torch.Tensor, values_in: torch.Tensor,
                          weights_input: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """Fused 1D conservative regrid: (S, n), (S, m), (S, n-1) -> (S, m-1)."""
    L = _lib.load()
    device = x_in.device
    S, n = x_in.shape
    m = x_out.shape[1]
    if values_in.shape != (S, n - 1):
        raise ValueError(f"values_in must have shape {(S, n - 1)}, got {tuple(values_in.shape)}")
    if out is None:
        out = torch.empty((S, m - 1), dtype=F64, device=device)
    with torch.cuda.device(device):
        _lib.check(L.rg_regrid1d_conservative(device.index, _stream(device), S, n, m, x_in.data_ptr(),
                                              x_out.data_ptr(), _lib.ptr(weights_input), values_in.data_ptr(),
                                              out.data_ptr()), "rg_regrid1d_conservative")
    return out


def find_indices_1d(x_in: torch.Tensor, x_out: torch.Tensor, fill_value: int, method: str) -> torch.Tensor:
    L = _lib.load()
    device = x_in.device
    D, n = x_in.shape
    m = x_out.shape[1]
    out = torch.empty((D, m), dtype=I64, device=device)
    code = {"brute": 0, "searchsorted": 1}[method]
    with torch.cuda.device(device):
        _lib.check(L.rg_find_indices_1d(device.index, _stream(device), code, D, n, m, x_in.data_ptr(),
                                        x_out.data_ptr(), int(fill_value), out.data_ptr()), "rg_find_indices_1d")
    return out


def find_indices_2d(x: torch.Tensor, y: torch.Tensor, px: torch.Tensor, py: torch.Tensor, fill_value: int) -> torch.Tensor:
    """Flat cell index i*(ny-1)+j of the lowest-index cell containing each point, else `fill_value`."""
    L = _lib.load()
    device = x.device
    nx, ny = x.shape
    npts = int(px.numel())
    out = torch.empty(px.shape, dtype=I64, device=device)
    with torch.cuda.device(device):
        nbytes = ctypes.c_size_t()
        _lib.check(L.rg_find_indices_2d_workspace_bytes(nx, ny, npts, ctypes.byref(nbytes)),
                   "rg_find_indices_2d_workspace_bytes")
        ws = _workspace(nbytes.value, device)
        _lib.check(L.rg_find_indices_2d(device.index, _stream(device), nx, ny, x.data_ptr(), y.data_ptr(), npts,
                                        px.data_ptr(), py.data_ptr(), int(fill_value), out.data_ptr(),
                                        ws.data_ptr(), ws.numel()), "rg_find_indices_2d")
    return out


BOUNDS_MODES = {"extrapolate": 0, "nan": 1, "raise": 2}


def multilinear1d_weights(x_in: torch.Tensor, x_out: torch.Tensor, weights_input: torch.Tensor | None = None,
                          bounds: str = "extrapolate"):
    """1D multilinear weights of D stacked grids (``rg_multilinear1d_weights``): ``(ii, io, v)`` of shape
    ``(D, 2 m)``, every row sorted by (input, output) like ``weights()`` returns it, and the number of output points
    outside their grid (only counted -- one host sync -- for ``bounds="raise"``)."""
    L = _lib.load()
    if bounds not in BOUNDS_MODES:
        raise ValueError(f"Unrecognized {bounds=}, expected one of ('extrapolate', 'nan', 'raise').")
    device = x_in.device
    D, n = x_in.shape
    m = x_out.shape[1]
    ii = torch.empty((D, 2 * m), dtype=I64, device=device)
    io = torch.empty((D, 2 * m), dtype=I64, device=device)
    v = torch.empty((D, 2 * m), dtype=F64, device=device)
    n_outside = 0
    # at most 65535 grids and 2^31 triplets per call
    per = max(1, min(65535, (2 ** 31 - 2) // max(2 * m, 1)))
    with torch.cuda.device(device):
        for d0 in range(0, D, per):
            d1 = min(D, d0 + per)
            nbytes = ctypes.c_size_t()
            _lib.check(L.rg_multilinear1d_workspace_bytes(d1 - d0, m, ctypes.byref(nbytes)), "rg_multilinear1d_workspace_bytes")
            ws = _workspace(nbytes.value, device)
            cnt = ctypes.c_int64()
            _lib.check(L.rg_multilinear1d_weights(device.index, _stream(device), d1 - d0, n, m, x_in[d0:d1].data_ptr(),
                                                  x_out[d0:d1].data_ptr(),
                                                  None if weights_input is None else weights_input[d0:d1].data_ptr(),
                                                  BOUNDS_MODES[bounds], ii[d0:d1].data_ptr(), io[d0:d1].data_ptr(),
                                                  v[d0:d1].data_ptr(), ctypes.byref(cnt) if bounds == "raise" else None,
                                                  ws.data_ptr(), ws.numel()), "rg_multilinear1d_weights")
            n_outside += int(cnt.value)
    return ii, io, v, n_outside


def sort_triplets(ii: torch.Tensor, io: torch.Tensor, v: torch.Tensor, n_in: int, n_out: int):
    """Raw triplets of one element -> sorted by (input, output), stable (``rg_sort_triplets``: the ordering of
    ``_coalesce``, _weights_arrays.py:54-59)."""
    L = _lib.load()
    device = v.device
    n = int(v.numel())
    oi, oo, ov = torch.empty_like(ii), torch.empty_like(io), torch.empty_like(v)
    if n == 0:
        return oi, oo, ov
    with torch.cuda.device(device):
        nbytes = ctypes.c_size_t()
        _lib.check(L.rg_sort_triplets_workspace_bytes(n, ctypes.byref(nbytes)), "rg_sort_triplets_workspace_bytes")
        ws = _workspace(nbytes.value, device)
        _lib.check(L.rg_sort_triplets(device.index, _stream(device), n, int(n_in), int(n_out), ii.data_ptr(), io.data_ptr(),
                                      v.data_ptr(), oi.data_ptr(), oo.data_ptr(), ov.data_ptr(), ws.data_ptr(), ws.numel()),
                   "rg_sort_triplets")
    return oi, oo, ov


def multilinear2d_weights(x: torch.Tensor, y: torch.Tensor, px: torch.Tensor, py: torch.Tensor,
                          bounds: str = "extrapolate") -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Bilinear weights of the points (px, py) on the curvilinear VERTEX grid (x, y): cell location
    (``rg_find_indices_2d``) + inverse bilinear map (``rg_multilinear2d_weights``).
    Returns ``idx4`` int64 [P, 4] (flat vertex indices, ascending), ``w4`` float64 [P, 4] and the device
    counter of points outside the grid (int32 [1]; reading it synchronises)."""
    L = _lib.load()
    if bounds not in BOUNDS_MODES:
        raise ValueError(f"Unrecognized {bounds=}, expected one of ('extrapolate', 'nan', 'raise').")
    device = x.device
    P = int(px.numel())
    px, py = px.reshape(-1).contiguous(), py.reshape(-1).contiguous()
    cell = find_indices_2d(x, y, px, py, -1)
    idx4 = torch.empty((max(P, 1), 4), dtype=I64, device=device)[:P]
    w4 = torch.empty((max(P, 1), 4), dtype=F64, device=device)[:P]
    n_out = torch.zeros(1, dtype=I32, device=device)
    nx, ny = x.shape
    with torch.cuda.device(device):
        _lib.check(L.rg_multilinear2d_weights(device.index, _stream(device), nx, ny, x.data_ptr(), y.data_ptr(), P,
                                              px.data_ptr(), py.data_ptr(), cell.data_ptr(), -1, BOUNDS_MODES[bounds],
                                              idx4.data_ptr(), w4.data_ptr(), n_out.data_ptr()),
                   "rg_multilinear2d_weights")
    return idx4, w4, n_out


def ell4_apply(idx4: torch.Tensor, w4: torch.Tensor, values_in: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """values_in (F, n_vertices) -> (F, P): four weighted vertices per output point (``rg_ell4_apply``)."""
    L = _lib.load()
    device = values_in.device
    F, n_in = values_in.shape
    P = int(idx4.shape[0])
    if out is None:
        out = torch.empty((F, P), dtype=F64, device=device)
    with torch.cuda.device(device):
        _lib.check(L.rg_ell4_apply(device.index, _stream(device), F, n_in, P, idx4.data_ptr(), w4.data_ptr(),
                                   values_in.data_ptr(), out.data_ptr()), "rg_ell4_apply")
    return out
