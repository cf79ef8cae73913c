"""
``weights()``: the reference's public entry point (``regridding/_weights/_weights.py:13-196``)
with the compiled kernels replaced by the CUDA library.

Return value (the saved-weights layout, ``_weights.py:31-36``): ``(weights, shape_input,
shape_output)`` where ``weights`` is an object ndarray of shape ``shape_orthogonal`` whose
elements are tuples ``(indices_input int64[nnz], indices_output int64[nnz], values
float64[nnz])`` sorted by (input, output) with unique pairs, and the two shapes are the
full CELL shapes including orthogonal axes.  It pickles exactly like the reference's.

Every element also stays resident on the GPU (see ``_cache``), so a following
``regrid_from_weights`` does not upload it again.
"""

from __future__ import annotations

from typing import Literal, Sequence

import numpy as np
import torch

from . import _cache, _device, _util

__all__ = ["weights", "weights_packed"]


def weights(
    coordinates_input,
    coordinates_output,
    axis_input: None | int | Sequence[int] = None,
    axis_output: None | int | Sequence[int] = None,
    weights_input=None,
    method: Literal["multilinear", "conservative"] = "multilinear",
    bounds: Literal["extrapolate", "nan", "raise"] = "extrapolate",
    perturb: None | bool = None,
    seed: "None | int | np.random.Generator" = _util.SEED_DEFAULT,
) -> tuple[np.ndarray, tuple[int, ...], tuple[int, ...]]:
    """Drop-in for ``regridding.weights`` (same arguments, same result layout)."""
    unit_weights = getattr(weights_input, "unit", None)
    if unit_weights is not None:
        weights_input = getattr(weights_input, "value")

    if method == "multilinear":
        from ._multilinear import weights_multilinear

        elements, shape_in, shape_out, shape_orth = weights_multilinear(
            coordinates_input, coordinates_output, axis_input, axis_output, weights_input, bounds, perturb, seed)
    elif method == "conservative":
        elements, shape_in, shape_out, shape_orth = _weights_conservative_device(
            coordinates_input, coordinates_output, axis_input, axis_output, weights_input, perturb, seed)
    else:
        raise ValueError(f"unrecognized method '{method}'")

    result = np.empty(len(elements), dtype=object)
    for k, dw in enumerate(elements):
        triple = dw.to_host()
        if unit_weights is not None:
            triple = (triple[0], triple[1], triple[2] << unit_weights)
        elif isinstance(dw, _device.HostWeights):
            pass  # built in chunks and already downloaded: no device copy is kept
        else:
            _cache.freeze(triple)  # in-place edits would leave the cached device copy behind: they raise instead
            _cache.remember(triple, dw)
        result[k] = triple
    return result.reshape(shape_orth), shape_in, shape_out


def weights_packed(
    coordinates_input,
    coordinates_output,
    axis_input: None | int | Sequence[int] = None,
    axis_output: None | int | Sequence[int] = None,
    weights_input=None,
    method: Literal["multilinear", "conservative"] = "multilinear",
    bounds: Literal["extrapolate", "nan", "raise"] = "extrapolate",
    perturb: None | bool = None,
    seed: "None | int | np.random.Generator" = _util.SEED_DEFAULT,
):
    """Same arguments and same numbers as ``weights()``, returned as ``PackedWeights`` (four flat arrays instead
    of one Python tuple per orthogonal slice; ``.to_reference()`` gives the reference layout, ``.save()`` /
    ``PackedWeights.load()`` a memory-mappable file).  SURVEY section 8 row f2: with one grid per spectrum or per
    frame (BASELINE configs 2 and 4) the tuples, not the kernels, are the cost of the reference layout."""
    from ._packed import pack_elements

    if getattr(weights_input, "unit", None) is not None:
        raise ValueError("weights_packed does not carry units; pass plain arrays")
    if method == "multilinear":
        from ._multilinear import weights_multilinear

        elements, shape_in, shape_out, shape_orth = weights_multilinear(
            coordinates_input, coordinates_output, axis_input, axis_output, weights_input, bounds, perturb, seed)
    elif method == "conservative":
        elements, shape_in, shape_out, shape_orth = _weights_conservative_device(
            coordinates_input, coordinates_output, axis_input, axis_output, weights_input, perturb, seed)
    else:
        raise ValueError(f"unrecognized method '{method}'")
    return pack_elements(elements, shape_in, shape_out, shape_orth)


def _weights_conservative_device(
    coordinates_input,
    coordinates_output,
    axis_input,
    axis_output,
    weights_input,
    perturb,
    seed,
    device=None,
) -> tuple[list[_device.DeviceWeights], tuple[int, ...], tuple[int, ...], tuple[int, ...]]:
    """``_weights_conservative`` (regridding/_weights/_weights_conservative.py:11-146) with
    the per-slice kernels on the GPU.  Returns one DeviceWeights per orthogonal slice
    (C order), the cell shapes and the orthogonal shape."""
    if perturb is None:  # wcons.py:21-25: jitter by default only for >= 2 coordinate arrays
        perturb = (not isinstance(coordinates_input, np.ndarray)) and len(coordinates_input) > 1

    (coords_in, coords_out, axis_in, axis_out, shape_in, shape_out, shape_orth) = \
        _util.normalize_input_output_coordinates(coordinates_input, coordinates_output, axis_input, axis_output,
                                                 perturb=perturb, seed=seed)
    coords_in = tuple(np.asarray(getattr(c, "value", c), dtype=np.float64) for c in coords_in)
    coords_out = tuple(np.asarray(getattr(c, "value", c), dtype=np.float64) for c in coords_out)

    shape_cells_in = tuple(s - 1 if (a - len(shape_in)) in axis_in else s for a, s in enumerate(shape_in))
    shape_cells_out = tuple(s - 1 if (a - len(shape_out)) in axis_out else s for a, s in enumerate(shape_out))

    if weights_input is not None:
        weights_input = np.broadcast_to(np.asarray(weights_input, dtype=np.float64), shape_cells_in)

    device = _device.cuda_device(device)
    ndim_resample = len(axis_in)
    n_slices = int(np.prod(shape_orth, dtype=np.int64))

    def stacked(c, axes):
        # resampled axes last, in ascending axis order; orthogonal axes flattened in C order
        src = tuple(sorted(axes))
        c = np.moveaxis(c, src, tuple(range(-len(src), 0)))
        return c.reshape((n_slices,) + c.shape[len(c.shape) - len(src):])

    if ndim_resample == 1:
        x_in, x_out = stacked(coords_in[0], axis_in), stacked(coords_out[0], axis_out)
        w = None if weights_input is None else stacked(weights_input, axis_in)
        elements = _conservative_1d(x_in, x_out, w, device)
    elif ndim_resample == 2:
        xi, yi = (stacked(c, axis_in) for c in coords_in)
        xo, yo = (stacked(c, axis_out) for c in coords_out)
        elements = []
        indices = list(np.ndindex(*shape_orth))
        if n_slices == 1:
            w = None if weights_input is None else weights_input[indices[0]]  # wcons.py:125-126: orthogonal axes lead
            elements.append(_device.build_weights_2d(xi[0], yi[0], xo[0], yo[0], w, device=device))
        else:
            # per-slice grids (BASELINE config 4): chunks of slices are uploaded, built back to back without host
            # synchronisation (rg_build2d_batched) and downloaded before the next chunk, so device memory does not
            # grow with the number of slices
            per = _SLICES_PER_CHUNK_2D
            for k0 in range(0, n_slices, per):
                ks = range(k0, min(n_slices, k0 + per))
                sl = [tuple(_device.to_device(a[k], device) for a in (xi, yi, xo, yo)) for k in ks]
                ws_in = None if weights_input is None else [weights_input[indices[k]] for k in ks]
                for dw in _device.build_weights_2d_batched(sl, ws_in, device=device):
                    elements.append(_device.HostWeights(*dw.to_host(), dw.n_in, dw.n_out))
    else:
        raise NotImplementedError("Regridding operations greater than 2D are not supported")  # wcons.py:141-144
    return elements, shape_cells_in, shape_cells_out, tuple(shape_orth)


_SLICES_PER_CHUNK_2D = 4    # 2D slices built per batched call (each keeps its triplets on the device until downloaded)
_CHUNK_BYTES_1D = 1 << 30  # device memory of one chunk of stacked 1D builds (3 arrays of (S, n + m) 8-byte entries)


def _conservative_1d(x_in: np.ndarray, x_out: np.ndarray, w, device) -> list:
    """1D conservative weights of S stacked spectra in the public layout (wcons.py:59-106 +
    warr.py:44-73).  The walk emits one triplet per overlap; for ascending grids that is
    already the (input, output)-sorted order, for descending ones the (negative,
    complemented) indices are re-sorted per spectrum on the device.

    The spectra are built in chunks bounded by ``_CHUNK_BYTES_1D`` of device memory and every chunk is
    downloaded (one copy per array) before the next is built: device memory does not grow with S, and there
    are no per-spectrum device copies.  A single spectrum keeps its device copy (so that
    ``regrid_from_weights(*weights(...))`` does not upload it again)."""
    S, n = x_in.shape
    m = x_out.shape[1]
    per = max(1, _CHUNK_BYTES_1D // (24 * (n + m)))
    elements: list = []
    for s0 in range(0, S, per):
        s1 = min(S, s0 + per)
        xi = _device.to_device(x_in[s0:s1], device)
        xo = _device.to_device(x_out[s0:s1], device)
        wd = None if w is None else _device.to_device(w[s0:s1], device)
        ii, io, v, counts = _device.cons1d_batched(xi, xo, wd)
        counts_h = counts.cpu().numpy()
        desc = ((xi[:, 0] >= xi[:, -1]) | (xo[:, 0] >= xo[:, -1])).cpu().numpy()
        if S == 1:
            c = int(counts_h[0])
            a, b, val = ii[0, :c], io[0, :c], v[0, :c]
            if desc[0] and c > 1:
                a, b, val = _sorted_1d(a, b, val)
            elements.append(_device.DeviceWeights(a.contiguous(), b.contiguous(), val.contiguous(), n - 1, m - 1))
            continue
        for k in np.flatnonzero(desc):  # rare: _coalesce's stable sort (warr.py:54-59) on the complemented indices
            c = int(counts_h[k])
            if c > 1:
                a, b, val = _sorted_1d(ii[k, :c], io[k, :c], v[k, :c])
                ii[k, :c], io[k, :c], v[k, :c] = a, b, val
        ii_h, io_h, v_h = ii.cpu().numpy(), io.cpu().numpy(), v.cpu().numpy()
        for k in range(s1 - s0):
            c = int(counts_h[k])
            # the saved layout keeps the reference's (possibly negative) indices
            elements.append(_device.HostWeights(ii_h[k, :c].copy(), io_h[k, :c].copy(), v_h[k, :c].copy(), n - 1, m - 1))
    return elements


def _sorted_1d(a, b, val):
    key = (a - a.min()) * (b.max() - b.min() + 1) + (b - b.min())
    order = torch.sort(key, stable=True).indices
    return a[order], b[order], val[order]
