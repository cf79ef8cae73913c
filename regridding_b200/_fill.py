"""
``fill()``: the reference's NaN / masked-cell repair (``regridding/_fill/_fill.py:10-109``,
``regridding/_fill/_gauss_seidel.py:13-139``) with the relaxation on the GPU (the step
before regridding in real pipelines).

Host side (as in the reference): mask (NaNs by default), starting guess (median of the valid cells along the
interpolation axes; NumPy's ``nanmedian`` itself, because its two-middle-values average decides the bits), axis
bookkeeping.  Device side: ``rg_fill_gauss_seidel_2d`` -- all iterations of all frames in one cooperative launch,
bit-identical to the reference's sequential red-black sweep (``csrc/rg_fill.cu``).
"""

from __future__ import annotations

import ctypes
import warnings
from typing import Literal, Sequence

import numpy as np
import torch

from . import _device, _lib

__all__ = ["fill"]


def fill(a, where=None, axis: None | int | Sequence[int] = None, method: Literal["gauss_seidel"] = "gauss_seidel",
         **kwargs):
    """Drop-in for ``regridding.fill``."""
    if where is None:
        where = np.isnan(a)
    if method == "gauss_seidel":
        return fill_gauss_seidel(a=a, where=where, axis=axis, **kwargs)
    raise ValueError(f"Unrecognized method '{method}'")


def _guess_median(a: np.ndarray, where: np.ndarray, axis: tuple[int, ...]) -> np.ndarray:
    masked = np.where(where, np.nan, a)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)  # slices without valid cells -> 0 below
        result = np.nanmedian(masked, axis=axis, keepdims=True)
    return np.where(np.isnan(result), 0, result)


def fill_gauss_seidel(a, where, axis, guess=None, num_iterations: int = 100) -> np.ndarray:
    unit = getattr(a, "unit", None)
    if unit is not None:
        a = a.value
    a = np.array(a, dtype=np.float64, copy=True)
    a, where = np.broadcast_arrays(a, np.asarray(where, dtype=bool))
    a = a.copy()
    nd = a.ndim
    ax = tuple(range(nd)) if axis is None else tuple(np.lib.array_utils.normalize_axis_tuple(axis, nd))
    if guess is None:
        guess = _guess_median(a, where, ax)
    a[where] = np.broadcast_to(getattr(guess, "value", guess), a.shape)[where]
    if len(ax) != 2:
        raise ValueError(f"The number of interpolation axes, {len(ax)},is not supported")
    last = (-2, -1)
    am = np.moveaxis(a, ax, last)  # first listed axis -> y (slow), second -> x (fast), _gauss_seidel.py:33-39
    wm = np.moveaxis(where, ax, last)
    shape_moved = am.shape
    ny, nx = shape_moved[-2:]
    a3 = np.ascontiguousarray(am.reshape(-1, ny, nx))
    w3 = np.ascontiguousarray(wm.reshape(-1, ny, nx))
    device = _device.cuda_device()
    out = np.empty_like(a3)
    per = max(1, min(a3.shape[0], (2**31 - 1) // (ny * nx))) if a3.shape[0] else 1
    for t0 in range(0, a3.shape[0], per):
        t1 = min(a3.shape[0], t0 + per)
        ad = _device.to_device(a3[t0:t1], device)
        gauss_seidel_2d_(ad, torch.from_numpy(w3[t0:t1]).to(device), num_iterations)
        torch.from_numpy(out[t0:t1]).copy_(ad)
    result = np.moveaxis(out.reshape(shape_moved), last, ax)
    if unit is None:
        return result
    return result << unit


def gauss_seidel_2d_(a: torch.Tensor, where: torch.Tensor, num_iterations: int) -> torch.Tensor:
    """In place on the device: ``a`` float64 (T, ny, nx) with the guess already stored, ``where`` bool (T, ny, nx)."""
    L = _lib.load()
    device = a.device
    T, ny, nx = a.shape
    idx = where.reshape(-1).nonzero().reshape(-1)  # flat indices of the missing cells, ascending
    r = idx % (ny * nx)
    j, i = r // nx, r % nx
    colour = (i + j) & 1
    level = ((i == nx - 1) & bool(nx & 1)).to(torch.int64) + ((j == ny - 1) & bool(ny & 1)).to(torch.int64)
    lists, counts = [], []
    for c in range(2):
        for lv in range(3):
            sel = idx[(colour == c) & (level == lv)].to(torch.int32).contiguous()
            lists.append(sel)
            counts.append(int(sel.numel()))
    ptrs = (ctypes.c_void_p * 6)(*[t.data_ptr() if t.numel() else None for t in lists])
    cnts = (ctypes.c_int64 * 6)(*counts)
    with torch.cuda.device(device):
        _lib.check(L.rg_fill_gauss_seidel_2d(device.index, _device._stream(device), a.data_ptr(), T, ny, nx, ptrs, cnts,
                                             int(num_iterations)), "rg_fill_gauss_seidel_2d")
    return a
