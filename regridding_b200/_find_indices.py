"""
``find_indices()``: the reference's public entry point
(``regridding/_find_indices/_find_indices.py:12-136``) on the GPU.

1D: both reference methods (``brute``: first cell with x[m] <= p <= x[m+1];
``searchsorted``: binary search with the reference's edge fix-ups).  2D is NEW -- the
reference raises ``ValueError`` for it -- and follows its internal locators
(``index_of_point_brute`` / ``index_of_point_secant``): the lowest-index cell whose quad
contains the point, ``fill_value`` on both axes if none does.
"""

from __future__ import annotations

from typing import Literal, Sequence

import numpy as np
import torch

from . import _device, _util

__all__ = ["find_indices"]


def find_indices(
    coordinates_input,
    coordinates_output,
    axis_input: None | int | Sequence[int] = None,
    axis_output: None | int | Sequence[int] = None,
    fill_value: None | int = None,
    method: Literal["brute", "searchsorted"] = "brute",
) -> tuple[np.ndarray, ...]:
    """Drop-in for ``regridding.find_indices`` (plus 2D grids)."""
    if method not in ("brute", "searchsorted"):
        raise ValueError(f"method `{method}` not recognized.")

    (coords_in, coords_out, axis_in, axis_out, shape_in, shape_out, shape_orth) = \
        _util.normalize_input_output_coordinates(coordinates_input, coordinates_output, axis_input, axis_output)

    if fill_value is None:
        fill_value = np.iinfo(int).max

    ndim = len(axis_in)
    if ndim > 2:
        raise ValueError(f"{ndim}-dimensional {'brute-force search' if method == 'brute' else 'searchsorted'} not supported")

    device = _device.cuda_device()
    D = int(np.prod(shape_orth, dtype=np.int64))
    src_in, src_out = tuple(sorted(axis_in)), tuple(sorted(axis_out))
    last = tuple(range(-ndim, 0))
    grid_in = tuple(shape_in[a] for a in src_in)
    grid_out = tuple(shape_out[a] for a in src_out)

    def stack(c, src, grid):
        c = np.asarray(getattr(c, "value", c), dtype=np.float64)
        return _device.to_device(np.moveaxis(c, src, last).reshape(D, *grid), device)

    cin = [stack(c, src_in, grid_in) for c in coords_in]
    cout = [stack(c, src_out, grid_out) for c in coords_out]

    if ndim == 1:
        idx = _device.find_indices_1d(cin[0], cout[0], fill_value, method)
        results = (idx,)
    else:
        ncy = grid_in[1] - 1
        res_i = torch.empty((D, *grid_out), dtype=torch.int64, device=device)
        res_j = torch.empty((D, *grid_out), dtype=torch.int64, device=device)
        for d in range(D):
            flat = _device.find_indices_2d(cin[0][d], cin[1][d], cout[0][d].reshape(-1), cout[1][d].reshape(-1), -1)
            flat = flat.reshape(grid_out)
            inside = flat >= 0
            res_i[d] = torch.where(inside, flat // ncy, fill_value)
            res_j[d] = torch.where(inside, flat % ncy, fill_value)
        results = (res_i, res_j)

    out = []
    for r in results:
        r = r.cpu().numpy().reshape(*shape_orth, *grid_out)
        out.append(np.moveaxis(r, last, src_out))
    return tuple(out)
