"""
``regrid_from_weights()`` and ``regrid()``: the reference's public entry points
(``regridding/_regrid/_regrid_from_weights.py:12-162``, ``regridding/_regrid/_regrid.py:12-149``)
with the scatter-add kernel (``rfw.py:165-182``) replaced by the CSR apply on the GPU.

NumPy in -> NumPy out (host <-> device copies included); CUDA tensors in -> CUDA tensor
out with no host round trip.
"""

from __future__ import annotations

from typing import Literal, Sequence

import numpy as np
import torch

from . import _cache, _device, _util

__all__ = ["regrid_from_weights", "regrid"]


def _orthogonal_shape(shape: tuple[int, ...], axis: tuple[int, ...]) -> tuple[int, ...]:
    return tuple(shape[a] for a in range(-len(shape), 0) if a not in axis)


def _device_weights_of(element, n_in: int, n_out: int, device) -> _device.DeviceWeights:
    if isinstance(element, _device.DeviceWeights):
        return element
    if isinstance(element, _device.HostWeights):
        element = element.to_host()
    indices_input, indices_output, values = element
    values = getattr(values, "value", values)  # unit-carrying values (rfw.py:134-141)
    key = (indices_input, indices_output, values)
    dw = _cache.lookup(key, device)
    if dw is None or dw.n_in != n_in or dw.n_out != n_out:
        dw = _device.DeviceWeights.from_host(indices_input, indices_output, values, n_in, n_out, device)
        _cache.remember(key, dw)
    return dw


_CHUNK_BYTES = 256 << 20  # host <-> device pipeline granularity
_RING = 3


def _host_pipeline(groups, apply, vin_h: np.ndarray, out_h: np.ndarray, device) -> None:
    """values on the HOST -> result on the HOST, as a 3-stage pipeline over chunks of orthogonal
    slices: H2D copy, apply and D2H copy run on three streams with a ring of device buffers, so the
    PCIe link carries input and output at the same time (full duplex) while the kernels run.
    Device memory use is bounded by the ring, not by the number of slices.  Every output cell
    is written by the apply (cells no weight reaches get +0.0), so no zero fill is needed."""
    import warnings

    D, n_in = vin_h.shape
    n_out = out_h.shape[1]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)  # read-only views of broadcast inputs are only read
        src_t = torch.from_numpy(vin_h)
    dst_t = torch.from_numpy(out_h)
    per = max(1, min(D, _CHUNK_BYTES // (8 * max(n_in, n_out, 1))))
    items = []
    for d, e, dw in groups:
        for a in range(d, e, per):
            items.append((a, min(e, a + per), dw))
    if not items:
        return
    ring = min(_RING, len(items))
    main = torch.cuda.current_stream(device)
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    buf_in = [torch.empty((per, n_in), dtype=torch.float64, device=device) for _ in range(ring)]
    buf_out = [torch.empty((per, n_out), dtype=torch.float64, device=device) for _ in range(ring)]
    applied = [None] * ring   # apply of the chunk that last used slot k finished (input buffer reusable)
    drained = [None] * ring   # D2H of the chunk that last used slot k finished (output buffer reusable)
    s_in.wait_stream(main)
    s_out.wait_stream(main)
    for k, (a, b, dw) in enumerate(items):
        slot = k % ring
        n = b - a
        with torch.cuda.stream(s_in):
            if applied[slot] is not None:
                s_in.wait_event(applied[slot])
            buf_in[slot][:n].copy_(src_t[a:b], non_blocking=True)
            loaded = s_in.record_event()
        main.wait_event(loaded)
        if drained[slot] is not None:
            main.wait_event(drained[slot])
        apply(dw, buf_in[slot][:n], buf_out[slot][:n])
        applied[slot] = main.record_event()
        with torch.cuda.stream(s_out):
            s_out.wait_event(applied[slot])
            dst_t[a:b].copy_(buf_out[slot][:n], non_blocking=True)
            drained[slot] = s_out.record_event()
    main.wait_stream(s_out)
    main.wait_stream(s_in)
    torch.cuda.synchronize(device)  # the ring buffers are idle from here on; results are on the host


def regrid_from_weights(
    weights,
    shape_input: tuple[int, ...],
    shape_output: tuple[int, ...],
    values_input,
    values_output=None,
    axis_input: None | int | Sequence[int] = None,
    axis_output: None | int | Sequence[int] = None,
):
    """Drop-in for ``regridding.regrid_from_weights``; ``weights`` may also be a ``PackedWeights``."""
    from ._packed import PackedWeights

    if isinstance(weights, PackedWeights):
        weights = weights.to_reference()[0]
    on_device = isinstance(values_input, torch.Tensor) and values_input.is_cuda
    unit = getattr(values_input, "unit", None)
    if unit is not None:
        values_input = values_input.value

    # axes are normalised against the WEIGHTS' shapes (rfw.py:64-68)
    axis_in = _util.normalize_axis(axis_input, len(shape_input))
    axis_out = _util.normalize_axis(axis_output, len(shape_output))

    orth_in = _orthogonal_shape(tuple(shape_input), axis_in)
    orth_out = _orthogonal_shape(tuple(shape_output), axis_out)
    vin_shape = tuple(values_input.shape) if np.ndim(values_input) > 0 else ()
    orth_val = _orthogonal_shape(vin_shape, axis_in) if vin_shape else ()
    shape_orth = np.broadcast_shapes(orth_in, orth_out, orth_val)

    axis_in = tuple(sorted(axis_in))
    axis_out = tuple(sorted(axis_out))
    full_in = _util._embed(shape_orth, axis_in, {a: shape_input[a] for a in axis_in})
    full_out = _util._embed(shape_orth, axis_out, {a: shape_output[a] for a in axis_out})

    weights_arr = np.broadcast_to(np.array(weights), shape_orth, subok=True)
    flat_weights = weights_arr.reshape(-1)
    unit_weights = getattr(flat_weights[0][2], "unit", None) if flat_weights.size and not isinstance(
        flat_weights[0], (_device.DeviceWeights, _device.HostWeights)) else None

    cells_in = tuple(full_in[a] for a in axis_in)
    cells_out = tuple(full_out[a] for a in axis_out)
    n_in = int(np.prod(cells_in, dtype=np.int64))
    n_out = int(np.prod(cells_out, dtype=np.int64))
    D = int(np.prod(shape_orth, dtype=np.int64))
    last_in = tuple(range(-len(axis_in), 0))
    last_out = tuple(range(-len(axis_out), 0))

    if on_device:
        device = values_input.device
        vin = torch.broadcast_to(values_input.to(torch.float64), full_in)
        vin = torch.movedim(vin, axis_in, last_in).reshape(D, n_in).contiguous()
        if values_output is not None and tuple(values_output.shape) != tuple(full_out):
            raise ValueError(f"{values_output.shape=} should be equal to {full_out}")
    else:
        device = _device.cuda_device()
        vin_h = np.broadcast_to(np.asarray(values_input, dtype=np.float64), full_in)
        vin_h = np.ascontiguousarray(np.moveaxis(vin_h, axis_in, last_in).reshape(D, n_in))
        if values_output is None:
            values_output = np.zeros(full_out, dtype=float)
        elif values_output.shape != full_out:
            raise ValueError(f"{values_output.shape=} should be equal to {full_out}")

    # consecutive orthogonal slices that share one weights element are applied in one launch
    def groups():
        d = 0
        while d < D:
            e = d + 1
            while e < D and flat_weights[e] is flat_weights[d]:
                e += 1
            yield d, e, _device_weights_of(flat_weights[d], n_in, n_out, device)
            d = e

    def apply(dw, src, dst):
        if len(cells_in) == 2 and len(cells_out) == 2:
            _device.apply_planned(dw.plan(cells_in, cells_out), src, dst)
        else:
            _device.apply_csr(dw.csr(), src, dst)

    moved_out_shape = tuple(shape_orth) + cells_out
    if on_device:
        out = torch.empty((D, n_out), dtype=torch.float64, device=device)
        for d, e, dw in groups():
            apply(dw, vin[d:e], out[d:e])
        res = torch.movedim(out.reshape(moved_out_shape), last_out, axis_out)
        if values_output is not None:
            values_output.copy_(res)
            res = values_output
        return res

    # host: same in-place behaviour as the reference (rfw.py:120-154): the caller's buffer is
    # updated in place only if its moved/reshaped view is already C-contiguous; otherwise it is
    # left zeroed and the result is a fresh array
    moved = np.moveaxis(values_output, axis_out, last_out)
    moved_shape = moved.shape
    if moved.flags.c_contiguous:
        target = moved.reshape(D, *cells_out)
    else:
        values_output.fill(0)
        target = np.empty((D, *cells_out), dtype=float)
    _host_pipeline(groups(), apply, vin_h, target.reshape(D, n_out), device)
    result = np.moveaxis(target.reshape(moved_shape), last_out, axis_out)

    if unit_weights is not None:
        unit = unit_weights if unit is None else unit * unit_weights
    if unit is None:
        return result
    return result << unit


def regrid(
    coordinates_input,
    coordinates_output,
    values_input,
    values_output=None,
    axis_input: None | int | Sequence[int] = None,
    axis_output: None | int | Sequence[int] = None,
    method: Literal["multilinear", "conservative"] = "multilinear",
    bounds: Literal["extrapolate", "nan", "raise"] = "extrapolate",
    perturb: None | bool = None,
    seed: "None | int | np.random.Generator" = _util.SEED_DEFAULT,
):
    """Drop-in for ``regridding.regrid`` (= ``weights`` then ``regrid_from_weights``,
    regridding/_regrid/_regrid.py:130-149).  For ``method="conservative"`` the weights stay
    on the GPU between the two steps."""
    from ._weights import _weights_conservative_device, weights

    n_coordinates = 1 if isinstance(coordinates_input, np.ndarray) else len(coordinates_input)
    if method == "conservative" and n_coordinates == 1 and not (
            isinstance(values_input, torch.Tensor) and values_input.is_cuda):
        return _regrid_conservative_1d_fused(coordinates_input, coordinates_output, values_input, values_output,
                                             axis_input, axis_output, perturb, seed)
    if method == "multilinear" and n_coordinates == 2 and all(
            np.ndim(c) == 2 for c in (*coordinates_input, *coordinates_output)):
        return _regrid_multilinear_2d_fused(coordinates_input, coordinates_output, values_input, values_output,
                                            axis_input, axis_output, bounds, perturb, seed)
    if method == "conservative":
        elements, shape_in, shape_out, shape_orth = _weights_conservative_device(
            coordinates_input, coordinates_output, axis_input, axis_output, None, perturb, seed)
        w = np.empty(len(elements), dtype=object)
        for k, dw in enumerate(elements):
            w[k] = dw
        w = w.reshape(shape_orth)
    else:
        w, shape_in, shape_out = weights(coordinates_input, coordinates_output, axis_input, axis_output,
                                         method=method, bounds=bounds, perturb=perturb, seed=seed)
    return regrid_from_weights(w, shape_in, shape_out, values_input, values_output, axis_input, axis_output)


_FUSED_1D_CHUNK_BYTES = 1 << 30


def _regrid_conservative_1d_fused(coordinates_input, coordinates_output, values_input, values_output,
                                  axis_input, axis_output, perturb, seed):
    """``regrid(method="conservative")`` along ONE axis without materialising weights: every
    orthogonal slice (spectrum) goes through the fused kernel ``rg_regrid1d_conservative``,
    which accumulates exactly like ``weights`` + ``regrid_from_weights`` would (same bits).
    This is the BASELINE.json config-2 path (1M spectra x 4096 bins with per-spectrum grids):
    the reference layout would need one Python tuple of three arrays per spectrum."""
    unit = getattr(values_input, "unit", None)
    if unit is not None:
        values_input = values_input.value
    if perturb is None:
        perturb = False  # wcons.py:21-25: jitter by default only for >= 2 coordinate arrays
    (coords_in, coords_out, axis_in, axis_out, shape_in, shape_out, orth_coords) = \
        _util.normalize_input_output_coordinates(coordinates_input, coordinates_output, axis_input, axis_output,
                                                 perturb=perturb, seed=seed)
    (a_in,), (a_out,) = axis_in, axis_out
    x_in = np.asarray(getattr(coords_in[0], "value", coords_in[0]), dtype=np.float64)
    x_out = np.asarray(getattr(coords_out[0], "value", coords_out[0]), dtype=np.float64)
    n, m = x_in.shape[a_in], x_out.shape[a_out]
    values_input = np.asarray(values_input, dtype=np.float64)
    orth_val = _orthogonal_shape(tuple(values_input.shape), (a_in,)) if values_input.ndim else ()
    shape_orth = np.broadcast_shapes(orth_coords, orth_val)
    full_in = _util._embed(shape_orth, (a_in,), {a_in: n - 1})
    full_out = _util._embed(shape_orth, (a_out,), {a_out: m - 1})
    D = int(np.prod(shape_orth, dtype=np.int64))

    def stacked(c, axis, size):
        c = np.broadcast_to(c, _util._embed(shape_orth, (axis,), {axis: size}))
        return np.moveaxis(c, axis, -1).reshape(D, size)

    xi_h, xo_h = stacked(x_in, a_in, n), stacked(x_out, a_out, m)
    vi_h = stacked(values_input, a_in, n - 1)
    if values_output is not None and values_output.shape != full_out:
        raise ValueError(f"{values_output.shape=} should be equal to {full_out}")

    device = _device.cuda_device()
    out_h = np.empty((D, m - 1), dtype=float)
    per = max(1, min(D, _FUSED_1D_CHUNK_BYTES // (8 * (2 * n + 2 * m))))
    for d in range(0, D, per):
        e = min(D, d + per)
        xi = _device.to_device(xi_h[d:e], device)
        xo = _device.to_device(xo_h[d:e], device)
        vi = _device.to_device(vi_h[d:e], device)
        res = _device.regrid1d_conservative(xi, xo, vi)
        torch.from_numpy(out_h[d:e]).copy_(res)

    last = (-1,)
    if values_output is None:
        result = np.moveaxis(out_h.reshape(tuple(shape_orth) + (m - 1,)), last, (a_out,))
    else:
        # the reference's in-place rule (rfw.py:120-154), see regrid_from_weights
        moved = np.moveaxis(values_output, (a_out,), last)
        if moved.flags.c_contiguous:
            moved.reshape(D, m - 1)[...] = out_h
            result = values_output
        else:
            values_output.fill(0)
            result = np.moveaxis(out_h.reshape(moved.shape), last, (a_out,))
    if unit is None:
        return result
    return result << unit


def _regrid_multilinear_2d_fused(coordinates_input, coordinates_output, values_input, values_output,
                                 axis_input, axis_output, bounds, perturb=None, seed=_util.SEED_DEFAULT):
    """``regrid(method="multilinear")`` between two 2D grids shared by every orthogonal slice of the values
    (BASELINE config 5): cell location, bilinear weights and the four-point gather all stay on the GPU and no
    triplets are materialised (``rg_find_indices_2d`` -> ``rg_multilinear2d_weights`` -> ``rg_ell4_apply``).
    Same numbers as ``weights(method="multilinear")`` + ``regrid_from_weights`` (same accumulation order)."""
    if bounds not in ("extrapolate", "nan", "raise"):
        raise ValueError(f"Unrecognized {bounds=}, expected one of ('extrapolate', 'nan', 'raise').")
    unit = getattr(values_input, "unit", None)
    if unit is not None:
        values_input = values_input.value
    (coords_in, coords_out, axis_in, axis_out, shape_in, shape_out, _orth) = \
        _util.normalize_input_output_coordinates(coordinates_input, coordinates_output, axis_input, axis_output,
                                                 perturb=bool(perturb), seed=seed)  # wml.py:58-59: None means False
    device = _device.cuda_device()
    x, y = (_device.to_device(np.asarray(getattr(c, "value", c), dtype=np.float64), device) for c in coords_in)
    px, py = (_device.to_device(np.asarray(getattr(c, "value", c), dtype=np.float64), device) for c in coords_out)
    grid_in, grid_out = tuple(x.shape), tuple(px.shape)
    on_device = isinstance(values_input, torch.Tensor) and values_input.is_cuda
    vals = values_input if on_device else np.asarray(values_input, dtype=np.float64)
    nd = vals.ndim
    # the axes returned by the normalisation are negative and relative to the COORDINATE arrays, exactly as
    # regrid_from_weights takes them against the weights' shapes (rfw.py:64-68): with bare 2D grids the resampled
    # axes of the values are always their last two, whatever positive axis numbers the caller wrote
    a_in = tuple(sorted(a % nd for a in axis_in))
    a_out = tuple(sorted(a % nd for a in axis_out))
    if tuple(vals.shape[a] for a in a_in) != grid_in:
        raise ValueError(f"values_input has shape {tuple(vals.shape)} along {a_in=}, expected the vertex grid {grid_in}")
    idx4, w4, n_outside = _device.multilinear2d_weights(x, y, px, py, bounds)
    if bounds == "raise":
        n_bad = int(n_outside.item())
        if n_bad:
            raise ValueError(f"{n_bad} of the output points fall outside the input grid, and {bounds=}.")
    last = (-2, -1)
    if on_device:
        moved = torch.movedim(vals.to(torch.float64), a_in, last)
        orth = tuple(moved.shape[:-2])
        F = int(np.prod(orth, dtype=np.int64))
        res = _device.ell4_apply(idx4, w4, moved.reshape(F, -1).contiguous())
        res = torch.movedim(res.reshape(*orth, *grid_out), last, a_out)
        if values_output is not None:
            values_output.copy_(res)
            return values_output
        return res
    moved = np.moveaxis(vals, a_in, last)
    orth = moved.shape[:-2]
    F = int(np.prod(orth, dtype=np.int64))
    flat = np.ascontiguousarray(moved.reshape(F, -1))
    out_h = np.empty((F, grid_out[0] * grid_out[1]), dtype=float)
    per = max(1, min(F, (1 << 30) // (8 * (flat.shape[1] + out_h.shape[1]))))
    for f in range(0, F, per):
        e = min(F, f + per)
        res = _device.ell4_apply(idx4, w4, _device.to_device(flat[f:e], device))
        torch.from_numpy(out_h[f:e]).copy_(res)
    full_out = np.moveaxis(out_h.reshape(*orth, *grid_out), last, a_out)
    if values_output is not None:
        if values_output.shape != full_out.shape:
            raise ValueError(f"{values_output.shape=} should be equal to {full_out.shape}")
        values_output[...] = full_out
        full_out = values_output
    if unit is None:
        return full_out
    return full_out << unit
