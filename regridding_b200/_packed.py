"""
Packed weights: the saved-weights layout of the reference without one Python tuple per slice.

The reference stores ``weights[0]`` as an object ndarray with one ``(indices_input, indices_output, values)``
tuple per orthogonal slice (``regridding/_weights/_weights.py:31-36``) and users pickle it.  With one grid per
spectrum (BASELINE config 2: 10^6 slices) or per frame (config 4) that is 3 x 10^6 small arrays: building,
pickling and re-uploading them is pure host overhead.  ``PackedWeights`` keeps the same numbers in four flat
arrays -- CSR over the slices -- and converts to / from the reference layout on demand (zero-copy views):

    offsets          int64 [D + 1]   slice d owns entries offsets[d] : offsets[d + 1]
    indices_input    int64 [T]       exactly the reference's values (flat cell indices, possibly negative
    indices_output   int64 [T]       for descending 1D grids), sorted by (input, output) inside a slice
    values           float64 [T]

On disk: one little-endian file -- magic, JSON header (shapes, counts, byte offsets), the four arrays at 64-byte
aligned offsets -- that ``load`` maps with ``numpy.memmap`` (no parsing, no copy; page-locked upload on demand).
"""

from __future__ import annotations

import json
import pathlib

import numpy as np

MAGIC = b"RGB200PW"
VERSION = 1
_ALIGN = 64
_FIELDS = (("offsets", np.int64), ("indices_input", np.int64), ("indices_output", np.int64), ("values", np.float64))


class PackedWeights:
    def __init__(self, offsets, indices_input, indices_output, values, shape_input, shape_output, shape_orthogonal):
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self.indices_input = np.asarray(indices_input, dtype=np.int64)
        self.indices_output = np.asarray(indices_output, dtype=np.int64)
        self.values = np.asarray(values, dtype=np.float64)
        self.shape_input = tuple(int(s) for s in shape_input)
        self.shape_output = tuple(int(s) for s in shape_output)
        self.shape_orthogonal = tuple(int(s) for s in shape_orthogonal)
        d = int(np.prod(self.shape_orthogonal, dtype=np.int64))
        if self.offsets.shape != (d + 1,):
            raise ValueError(f"offsets must have {d + 1} entries for {self.shape_orthogonal=}, got {self.offsets.shape}")
        t = int(self.offsets[-1]) if self.offsets.size else 0
        if self.offsets[0] != 0 or np.any(np.diff(self.offsets) < 0):
            raise ValueError("offsets must start at 0 and be non-decreasing")
        for name in ("indices_input", "indices_output", "values"):
            if getattr(self, name).shape != (t,):
                raise ValueError(f"{name} must have {t} entries, got {getattr(self, name).shape}")

    # ------------------------------------------------------------------ basic access
    def __len__(self) -> int:
        return self.offsets.size - 1

    @property
    def nnz(self) -> int:
        return int(self.offsets[-1])

    def element(self, d: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Slice ``d`` (C order over the orthogonal axes) as the reference's tuple; views, no copy."""
        a, b = int(self.offsets[d]), int(self.offsets[d + 1])
        return self.indices_input[a:b], self.indices_output[a:b], self.values[a:b]

    # ------------------------------------------------------------------ reference layout
    @classmethod
    def from_reference(cls, weights) -> "PackedWeights":
        """``weights`` = what ``regridding.weights`` returns: ``(object ndarray, shape_input, shape_output)``."""
        w, shape_in, shape_out = weights
        w = np.asarray(w, dtype=object)
        flat = w.reshape(-1)
        counts = np.fromiter((len(e[2]) for e in flat), dtype=np.int64, count=flat.size)
        offsets = np.zeros(flat.size + 1, dtype=np.int64)
        np.cumsum(counts, out=offsets[1:])
        t = int(offsets[-1])
        ii, io, v = np.empty(t, np.int64), np.empty(t, np.int64), np.empty(t, np.float64)
        for d, e in enumerate(flat):
            a, b = offsets[d], offsets[d + 1]
            ii[a:b], io[a:b], v[a:b] = e[0], e[1], getattr(e[2], "value", e[2])
        return cls(offsets, ii, io, v, shape_in, shape_out, w.shape)

    def to_reference(self):
        """``(weights, shape_input, shape_output)`` in the reference's layout; the tuples hold views."""
        out = np.empty(len(self), dtype=object)
        for d in range(len(self)):
            out[d] = self.element(d)
        return out.reshape(self.shape_orthogonal), self.shape_input, self.shape_output

    # ------------------------------------------------------------------ file format
    def save(self, path) -> None:
        path = pathlib.Path(path)
        arrays = [np.ascontiguousarray(getattr(self, n), dtype=dt).astype(np.dtype(dt).newbyteorder("<"), copy=False)
                  for n, dt in _FIELDS]
        header = {"version": VERSION, "shape_input": self.shape_input, "shape_output": self.shape_output,
                  "shape_orthogonal": self.shape_orthogonal, "arrays": {}}
        # two passes: the header length decides where the arrays start
        for _ in range(2):
            blob = json.dumps(header).encode()
            pos = (len(MAGIC) + 8 + len(blob) + 256 + _ALIGN - 1) // _ALIGN * _ALIGN  # slack for the second pass
            for (n, dt), a in zip(_FIELDS, arrays):
                header["arrays"][n] = {"dtype": np.dtype(dt).str.replace("=", "<"), "count": int(a.size), "offset": pos}
                pos = (pos + a.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        blob = json.dumps(header).encode()
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(np.uint64(len(blob)).tobytes())
            f.write(blob)
            for (n, _), a in zip(_FIELDS, arrays):
                f.seek(header["arrays"][n]["offset"])
                f.write(a.tobytes() if a.size else b"")
            f.truncate(max(pos, f.tell()))

    @classmethod
    def load(cls, path, mmap: bool = True) -> "PackedWeights":
        path = pathlib.Path(path)
        with open(path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise ValueError(f"{path} is not a packed weights file")
            n = int(np.frombuffer(f.read(8), dtype="<u8")[0])
            header = json.loads(f.read(n).decode())
        if header.get("version") != VERSION:
            raise ValueError(f"unsupported packed weights version {header.get('version')}")
        arrays = {}
        for name, _ in _FIELDS:
            meta = header["arrays"][name]
            if meta["count"] == 0:
                arrays[name] = np.empty(0, dtype=meta["dtype"])
            elif mmap:
                arrays[name] = np.memmap(path, dtype=meta["dtype"], mode="r", offset=meta["offset"], shape=(meta["count"],))
            else:
                arrays[name] = np.fromfile(path, dtype=meta["dtype"], count=meta["count"], offset=meta["offset"])
        return cls(arrays["offsets"], arrays["indices_input"], arrays["indices_output"], arrays["values"],
                   header["shape_input"], header["shape_output"], header["shape_orthogonal"])

    # ------------------------------------------------------------------ device
    def to_device(self, device=None) -> list:
        """One ``DeviceWeights`` per slice, all views of three device tensors uploaded in one copy each."""
        import torch

        from . import _device

        device = _device.cuda_device(device)
        n_in, n_out = _resampled_sizes(self)
        ii = _device.to_device(np.asarray(self.indices_input), device, _device.I64)
        io = _device.to_device(np.asarray(self.indices_output), device, _device.I64)
        v = _device.to_device(np.asarray(self.values), device, _device.F64)
        out = []
        for d in range(len(self)):
            a, b = int(self.offsets[d]), int(self.offsets[d + 1])
            out.append(_device.DeviceWeights(ii[a:b], io[a:b], v[a:b], n_in, n_out))
        del torch
        return out


def _resampled_sizes(p: PackedWeights) -> tuple[int, int]:
    """Cells per slice on the input / output side: the resampled axes are the ones that differ from the
    orthogonal shape... which the file does not record, so derive them from the totals."""
    d = max(1, int(np.prod(p.shape_orthogonal, dtype=np.int64)))
    n_in = int(np.prod(p.shape_input, dtype=np.int64)) // d
    n_out = int(np.prod(p.shape_output, dtype=np.int64)) // d
    return n_in, n_out


def pack_elements(elements, shape_input, shape_output, shape_orthogonal) -> PackedWeights:
    """Device elements (one ``DeviceWeights`` per slice) -> ``PackedWeights``: three concatenations on the
    device and one download per array instead of 3 x D small copies."""
    import torch

    counts = np.fromiter((e.nnz for e in elements), dtype=np.int64, count=len(elements))
    offsets = np.zeros(len(elements) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    if len(elements) == 0 or int(offsets[-1]) == 0:
        z = np.empty(0)
        return PackedWeights(offsets, z.astype(np.int64), z.astype(np.int64), z, shape_input, shape_output, shape_orthogonal)
    if all(isinstance(e.values, np.ndarray) for e in elements):  # built in chunks and already on the host
        ii = np.concatenate([e.indices_input for e in elements])
        io = np.concatenate([e.indices_output for e in elements])
        v = np.concatenate([e.values for e in elements])
        return PackedWeights(offsets, ii, io, v, shape_input, shape_output, shape_orthogonal)
    host = lambda t: t if isinstance(t, np.ndarray) else t.cpu().numpy()  # noqa: E731
    if any(isinstance(e.values, np.ndarray) for e in elements):
        ii = np.concatenate([host(e.indices_input) for e in elements])
        io = np.concatenate([host(e.indices_output) for e in elements])
        v = np.concatenate([host(e.values) for e in elements])
        return PackedWeights(offsets, ii, io, v, shape_input, shape_output, shape_orthogonal)
    ii = torch.cat([e.indices_input for e in elements]).cpu().numpy()
    io = torch.cat([e.indices_output for e in elements]).cpu().numpy()
    v = torch.cat([e.values for e in elements]).cpu().numpy()
    return PackedWeights(offsets, ii, io, v, shape_input, shape_output, shape_orthogonal)
