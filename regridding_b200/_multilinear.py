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

    # location (searchsorted + the reference's fix-ups), weights and the saved-weights ordering in one C-ABI call
    # (rg_multilinear1d_weights, csrc/rg_multilinear1d.cu)
    ii, io, vv, n_outside = _device.multilinear1d_weights(x_in, x_out, w_in, bounds)
    if bounds == "raise" and n_outside:
        raise ValueError(f"{n_outside} of the output points fall outside the input grid, and {bounds=}.")
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
        ii, io, vv = _device.sort_triplets(ii.contiguous(), io.contiguous(), vv.contiguous(), n_in, P)
        elements.append(_device.DeviceWeights(ii, io, vv, n_in, P))
    return elements, tuple(shape_in), tuple(shape_out), tuple(shape_orth)
