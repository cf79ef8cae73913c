"""
``ndarray_linear_interpolation()``: the reference's public entry point
(``regridding/_interp_ndarray.py:11-166``) on the GPU -- like ``scipy.ndimage.map_coordinates(order=1)`` but along
only some of the axes of ``a``, with linear extrapolation outside (the cell index is clamped, not the coordinate).

The axis bookkeeping below follows the reference (same ``ValueError`` / ``NotImplementedError`` behaviour); the
per-slice kernels ``_ndarray_linear_interpolation_1d`` / ``_2d`` are ``rg_interp_linear_1d`` / ``rg_interp_bilinear_2d``
(``csrc/rg_interp.cu``), bit-identical to the reference's compiled arithmetic (goldens in ``tests/golden/golden_v3.npz``).
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _device, _lib

__all__ = ["ndarray_linear_interpolation"]


def ndarray_linear_interpolation(a, indices, axis=None, axis_indices=None):
    """Drop-in for ``regridding.ndarray_linear_interpolation``."""
    a = np.asarray(a)
    shape_a, ndim_a = a.shape, a.ndim
    shape_indices = np.broadcast_shapes(*(np.shape(ind) for ind in indices))
    ndim_indices = len(shape_indices)

    if axis is None:
        axis = tuple(range(ndim_a))
    axis = np.lib.array_utils.normalize_axis_tuple(axis, ndim=ndim_a)
    axis = tuple(int(ax) - ndim_a for ax in axis)  # negative axes, as the reference keeps them
    if axis_indices is None:
        axis_indices = tuple(range(ndim_indices))
    axis_indices = np.lib.array_utils.normalize_axis_tuple(axis=axis_indices, ndim=ndim_indices)
    axis_indices = tuple(int(ax) - ndim_indices for ax in axis_indices)

    if len(indices) != len(axis):
        raise ValueError(
            f"The number of indices, {len(indices)}, must match the number of elements in axis, {len(axis)}"
        )
    if len(axis) not in (1, 2):
        raise NotImplementedError

    orth_a = tuple(ax for ax in range(-ndim_a, 0) if ax not in axis)
    orth_ind = tuple(ax for ax in range(-ndim_indices, 0) if ax not in axis_indices)
    shape_orth = np.broadcast_shapes(tuple(shape_a[ax] for ax in orth_a), tuple(shape_indices[ax] for ax in orth_ind))
    ndim_ba, ndim_bi = len(shape_orth) + len(axis), len(shape_orth) + len(axis_indices)
    shape_ba = tuple(shape_a[ax] if ax in axis else shape_orth[orth_a.index(ax)] for ax in range(-ndim_ba, 0))
    shape_bi = tuple(shape_indices[ax] if ax in axis_indices else shape_orth[orth_ind.index(ax)]
                     for ax in range(-ndim_bi, 0))

    # orthogonal axes first (C order), interpolation axes last -- one kernel launch for all the slices
    a_b = np.broadcast_to(np.asarray(a, dtype=np.float64), shape_ba)
    src_a = tuple(ax for ax in range(-ndim_ba, 0) if ax not in axis) + tuple(axis)
    a_m = np.ascontiguousarray(np.transpose(a_b, [ax % ndim_ba for ax in src_a]))
    grid = tuple(shape_a[ax] for ax in axis)
    D = int(np.prod(shape_orth, dtype=np.int64))
    src_i = tuple(ax for ax in range(-ndim_bi, 0) if ax not in axis_indices) + tuple(axis_indices)
    perm_i = [ax % ndim_bi for ax in src_i]
    ind_m = [np.ascontiguousarray(np.transpose(np.broadcast_to(np.asarray(ind, dtype=np.float64), shape_bi), perm_i))
             for ind in indices]
    shape_moved = ind_m[0].shape
    P = int(np.prod(shape_moved[len(shape_orth):], dtype=np.int64))

    device = _device.cuda_device()
    L = _lib.load()
    ad = _device.to_device(a_m.reshape(D, -1), device)
    xd = [_device.to_device(im.reshape(D, P), device) for im in ind_m]
    out = torch.empty((D, P), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        if len(axis) == 1:
            _lib.check(L.rg_interp_linear_1d(device.index, st, D, grid[0], P, grid[0], P, ad.data_ptr(), xd[0].data_ptr(),
                                             out.data_ptr()), "rg_interp_linear_1d")
        else:
            _lib.check(L.rg_interp_bilinear_2d(device.index, st, D, grid[0], grid[1], P, grid[0] * grid[1], P,
                                               ad.data_ptr(), xd[0].data_ptr(), xd[1].data_ptr(), out.data_ptr()),
                       "rg_interp_bilinear_2d")
    res = out.cpu().numpy().reshape(shape_moved)
    inv = np.argsort(perm_i)
    return np.ascontiguousarray(np.transpose(res, inv))
