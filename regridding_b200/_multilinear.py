"""
1D multilinear (linear-interpolation) weights on the GPU: the reference's
``_weights_multilinear`` (``regridding/_weights/_weights_multilinear.py:9-206``).
Adjacent to the conservative hot path (it is ``weights()``'s default method): cell
location runs in the CUDA library, the two weights per output point are elementwise
IEEE operations on device tensors.

2D (two coordinate arrays) is NEW: the reference raises (``wml.py:128-131``).  It extends the 1D rule to a
curvilinear vertex grid -- containing cell from the 2D ``find_indices`` walk, (u, v) of the cell's bilinear
map, four weights per output point (``csrc/rg_multilinear2d.cu``) -- and is what BASELINE config 5 names.
More than two axes raise like the reference.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _device, _util


def weights_multilinear(coordinates_input, coordinates_output, axis_input, axis_output, weights_input,
                        bounds, perturb, seed):
    if bounds not in ("extrapolate", "nan", "raise"):
        raise ValueError(f"Unrecognized {bounds=}, expected one of ('extrapolate', 'nan', 'raise').")
    if perturb is None:
        perturb = False
    (coords_in, coords_out, axis_in, axis_out, shape_in, shape_out, shape_orth) = \
        _util.normalize_input_output_coordinates(coordinates_input, coordinates_output, axis_input, axis_output,
                                                 perturb=perturb, seed=seed)
    if len(axis_in) == 2:
        return _weights_multilinear_2d(coords_in, coords_out, axis_in, axis_out, shape_in, shape_out, shape_orth,
                                       weights_input, bounds)
    if len(axis_in) != 1:
        raise ValueError(f"{len(axis_in)}-dimensional multilinear interpolation is not supported")

    device = _device.cuda_device()
    D = int(np.prod(shape_orth, dtype=np.int64))
    n, m = shape_in[axis_in[0]], shape_out[axis_out[0]]

    def stack(c, ax, size):
        c = np.asarray(getattr(c, "value", c), dtype=np.float64)
        return _device.to_device(np.moveaxis(c, ax, -1).reshape(D, size), device)

    x_in = stack(coords_in[0], axis_in[0], n)
    x_out = stack(coords_out[0], axis_out[0], m)
    w_in = None
    if weights_input is not None:
        w_in = stack(np.broadcast_to(weights_input, shape_in), axis_in[0], n)

    fill = np.iinfo(int).max
    index = _device.find_indices_1d(x_in, x_out, fill, "searchsorted")
    index_max = n - 2
    outside = (index < 0) | (index > index_max)
    if bounds == "raise" and bool(outside.any().item()):
        raise ValueError(f"{int(outside.sum().item())} of the output points fall outside the input grid, and {bounds=}.")
    below = x_out < x_in[:, :1]
    i0 = torch.where(below, torch.zeros_like(index), index.clamp(0, index_max))  # wml.py:105-119
    i1 = i0 + 1
    x0 = torch.gather(x_in, 1, i0)
    x1 = torch.gather(x_in, 1, i1)
    w1 = (x_out - x0) / (x1 - x0)  # wml.py:185-186
    w0 = 1 - w1
    if w_in is not None:
        w0 = w0 * torch.gather(w_in, 1, i0)
        w1 = w1 * torch.gather(w_in, 1, i1)
    if bounds == "nan":
        nan = torch.full_like(w0, float("nan"))
        w0 = torch.where(outside, nan, w0)
        w1 = torch.where(outside, nan, w1)

    i_out = torch.arange(m, device=device, dtype=torch.int64).expand(D, m)
    ii = torch.stack((i0, i1), dim=2).reshape(D, 2 * m)
    io = torch.stack((i_out, i_out), dim=2).reshape(D, 2 * m)
    vv = torch.stack((w0, w1), dim=2).reshape(D, 2 * m)
    # canonical layout (_weights_arrays.py:44-73): stable sort by (input, output); the pairs are unique
    key = ii * m + io
    order = torch.sort(key, dim=1, stable=True).indices
    ii, io, vv = torch.gather(ii, 1, order), torch.gather(io, 1, order), torch.gather(vv, 1, order)

    elements = [_device.DeviceWeights(ii[d].contiguous(), io[d].contiguous(), vv[d].contiguous(), n, m)
                for d in range(D)]
    return elements, tuple(shape_in), tuple(shape_out), tuple(shape_orth)


def _weights_multilinear_2d(coords_in, coords_out, axis_in, axis_out, shape_in, shape_out, shape_orth,
                            weights_input, bounds):
    """Four triplets per output point, in the saved-weights layout (sorted by (input, output))."""
    device = _device.cuda_device()
    D = int(np.prod(shape_orth, dtype=np.int64))
    src_in, src_out = tuple(sorted(axis_in)), tuple(sorted(axis_out))
    last = (-2, -1)
    grid_in = tuple(shape_in[a] for a in src_in)
    grid_out = tuple(shape_out[a] for a in src_out)

    def stack(c, src, grid):
        c = np.asarray(getattr(c, "value", c), dtype=np.float64)
        return np.moveaxis(c, src, last).reshape(D, *grid)

    xin, yin = (stack(c, src_in, grid_in) for c in coords_in)
    xout, yout = (stack(c, src_out, grid_out) for c in coords_out)
    w_in = None
    if weights_input is not None:
        w_in = stack(np.broadcast_to(weights_input, shape_in), src_in, grid_in)
    n_in = grid_in[0] * grid_in[1]
    P = grid_out[0] * grid_out[1]
    elements = []
    for d in range(D):
        x, y = _device.to_device(xin[d], device), _device.to_device(yin[d], device)
        px, py = _device.to_device(xout[d].reshape(-1), device), _device.to_device(yout[d].reshape(-1), device)
        idx4, w4, n_outside = _device.multilinear2d_weights(x, y, px, py, bounds)
        if bounds == "raise":
            n_bad = int(n_outside.item())
            if n_bad:
                raise ValueError(f"{n_bad} of the output points fall outside the input grid, and {bounds=}.")
        ii = idx4.reshape(-1)
        io = torch.arange(P, device=device, dtype=torch.int64).repeat_interleave(4)
        vv = w4.reshape(-1)
        if w_in is not None:
            vv = vv * _device.to_device(w_in[d].reshape(-1), device)[ii]
        # canonical layout (_weights_arrays.py:44-73): stable sort by (input, output); the pairs are unique
        order = torch.sort(ii * P + io, stable=True).indices
        elements.append(_device.DeviceWeights(ii[order].contiguous(), io[order].contiguous(), vv[order].contiguous(),
                                              n_in, P))
    return elements, tuple(shape_in), tuple(shape_out), tuple(shape_orth)
