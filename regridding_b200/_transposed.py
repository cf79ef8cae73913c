"""
``transpose_weights()`` and ``transpose_weights_conservative()``: the reference's entry points
(``regridding/_weights/_weights_transposed/_weights_transposed.py:13-52, 55-262``) for running a
saved transform in the opposite direction (the README's iterative-inversion use case).

The plain transpose only swaps the two index arrays of every element.  The conservative one
also rescales every weight by ``volume_input[i] / volume_output[j]`` (and by
``1 / weights_input[i]**2`` when the forward weights were built with ``weights_input``); the cell
volumes (``_cell_volume``, ``:265-340``) and the rescale run on the GPU, bit-identical to the
reference (``rg_grid_area`` / ``rg_cell_length_1d`` / ``rg_transpose_conservative``).
"""

from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _cache, _device, _util

__all__ = ["transpose_weights", "transpose_weights_conservative"]


def transpose_weights(weights):
    """Drop-in for ``regridding.transpose_weights``: ``(i, j, w) -> (j, i, w)`` without copying."""
    weights_array, shape_input, shape_output = weights
    shape = weights_array.shape
    flat = weights_array.reshape(-1)
    result = np.empty(flat.size, dtype=object)
    for d in range(flat.size):
        indices_input, indices_output, values = flat[d]
        result[d] = (indices_output, indices_input, values)
    return result.reshape(shape), shape_output, shape_input


def _cell_volume(coords: tuple[np.ndarray, ...], axis: tuple[int, ...], n_slices: int, device) -> torch.Tensor:
    """(n_slices, n_cells) cell volumes, cells flattened in ascending-axis order (wT.py:265-304)."""
    src = tuple(sorted(axis))
    last = tuple(range(-len(src), 0))
    moved = [np.moveaxis(np.asarray(getattr(c, "value", c), dtype=np.float64), src, last) for c in coords]
    if len(coords) == 1:
        x = _device.to_device(moved[0].reshape(n_slices, moved[0].shape[-1]), device)
        return _device.cell_length_1d(x)
    if len(coords) == 2:
        nx, ny = moved[0].shape[-2:]
        xs = moved[0].reshape(n_slices, nx, ny)
        ys = moved[1].reshape(n_slices, nx, ny)
        out = torch.empty((n_slices, (nx - 1) * (ny - 1)), dtype=torch.float64, device=device)
        for d in range(n_slices):
            out[d] = _device.grid_area(xs[d], ys[d], device=device).reshape(-1)
        return out
    raise ValueError("Grids greater than 2D not supported.")  # wT.py:301-302


def transpose_weights_conservative(
    weights,
    coordinates_input,
    coordinates_output,
    axis_input: None | int | Sequence[int] = None,
    axis_output: None | int | Sequence[int] = None,
    weights_input=None,
):
    """Drop-in for ``regridding.transpose_weights_conservative``."""
    weights_array, shape_input, shape_output = weights
    (coords_in, coords_out, axis_in, axis_out, _, _, shape_orth) = _util.normalize_input_output_coordinates(
        coordinates_input, coordinates_output, axis_input, axis_output)
    # the stored flat indices address cells in ascending-axis order (wT.py:208-215)
    axis_in = tuple(sorted(axis_in))
    axis_out = tuple(sorted(axis_out))
    n_slices = int(np.prod(shape_orth, dtype=np.int64))
    device = _device.cuda_device()

    w = None
    if weights_input is not None:
        w = np.broadcast_to(np.asarray(getattr(weights_input, "value", weights_input), dtype=np.float64), shape_input)
        w = np.moveaxis(w, axis_in, tuple(range(-len(axis_in), 0))).reshape(n_slices, -1)
        w = _device.to_device(w, device)
    vol_in = _cell_volume(coords_in, axis_in, n_slices, device)
    vol_out = _cell_volume(coords_out, axis_out, n_slices, device)

    shape = weights_array.shape
    flat = weights_array.reshape(-1)
    result = np.empty(flat.size, dtype=object)
    for d in range(flat.size):
        indices_input, indices_output, values = flat[d]
        dw = _cache.lookup((indices_input, indices_output, values), device)
        if dw is None or dw.n_in != vol_in.shape[1] or dw.n_out != vol_out.shape[1]:
            # range-checked upload (weights built for other grid shapes raise IndexError instead of reading out
            # of bounds on the device); the transposed values are computed from the wrapped indices
            dw = _device.DeviceWeights.from_host(indices_input, indices_output, values, vol_in.shape[1],
                                                 vol_out.shape[1], device)
        v_t = _device.transpose_conservative(dw, vol_in[d], vol_out[d], None if w is None else w[d])
        result[d] = (indices_output, indices_input, v_t.cpu().numpy())
    return result.reshape(shape), shape_output, shape_input
