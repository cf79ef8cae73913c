"""
Host-side argument normalisation shared by ``weights``, ``regrid_from_weights`` and
``find_indices``.

Mirrors the semantics of ``regridding/_util.py:12-139`` of the reference (axis
normalisation to negative indices, broadcasting of the coordinate tuples against the
orthogonal shape, the ValueErrors, and the seeded 1e-9 jitter of the OUTPUT
coordinates).  The jitter is drawn ON THE HOST with NumPy's own generator because the
reference does so (``_util.py:121-129``): PCG64 + ziggurat consumes a data-dependent
number of words, so any other generator would change every weight.
"""

from __future__ import annotations

from typing import Sequence

import numpy as np

SEED_DEFAULT = 42  # regridding/_util.py:4
EPSILON_PERTURB = 1e-9  # regridding/_util.py:122


def normalize_axis(axis: None | int | Sequence[int], ndim: int) -> tuple[int, ...]:
    """Axes as a tuple of NEGATIVE indices (all axes for ``None``); _util.py:12-20."""
    if axis is None:
        axis = tuple(range(ndim))
    axis = np.lib.array_utils.normalize_axis_tuple(axis, ndim=ndim)
    return tuple(int(a) - ndim for a in axis)


def _embed(shape_orthogonal: tuple[int, ...], axis: tuple[int, ...], sizes: dict[int, int]) -> tuple[int, ...]:
    """Full shape whose axes `axis` (negative) have the given sizes and whose remaining
    axes, in order, are `shape_orthogonal`."""
    ndim = len(shape_orthogonal) + len(axis)
    orth = iter(shape_orthogonal)
    return tuple(sizes[a] if a in sizes else next(orth) for a in range(-ndim, 0))


def normalize_input_output_coordinates(
    coordinates_input,
    coordinates_output,
    axis_input: None | int | Sequence[int] = None,
    axis_output: None | int | Sequence[int] = None,
    perturb: bool = False,
    seed: "None | int | np.random.Generator" = SEED_DEFAULT,
):
    """Returns ``(coordinates_input, coordinates_output, axis_input, axis_output,
    shape_coordinates_input, shape_coordinates_output, shape_orthogonal)`` with the
    coordinates broadcast to their full shapes and the axes negative, sorted descending."""
    if isinstance(coordinates_input, np.ndarray):
        coordinates_input = (coordinates_input,)
    if isinstance(coordinates_output, np.ndarray):
        coordinates_output = (coordinates_output,)

    # unit-carrying outputs pull the inputs to the same unit (duck typed, _util.py:45-54)
    converted = []
    for c_in, c_out in zip(coordinates_input, coordinates_output):
        unit = getattr(c_out, "unit", None)
        converted.append(c_in << unit if unit is not None else c_in)
    coordinates_input = tuple(converted)

    shape_coords_in = np.broadcast(*coordinates_input).shape
    shape_coords_out = np.broadcast(*coordinates_output).shape

    axis_input = tuple(sorted(normalize_axis(axis_input, len(shape_coords_in)), reverse=True))
    axis_output = tuple(sorted(normalize_axis(axis_output, len(shape_coords_out)), reverse=True))

    if len(axis_output) != len(axis_input):
        raise ValueError(
            f"The number of axes in `axis_output`, {axis_output}, "
            f"must match the number of axes in `axis_input`, {axis_input}"
        )
    if len(coordinates_input) != len(axis_input):
        raise ValueError(
            f"The number of elements in `coordinates_input`, {len(coordinates_input)}, "
            f"should match the number of axes in `axis_input`, {axis_input}"
        )
    if len(coordinates_output) != len(coordinates_input):
        raise ValueError(
            f"The number of elements in `coordinates_output`, {len(coordinates_output)}, "
            f"should match the number of elements in `coordinates_input`, {len(coordinates_input)}"
        )

    orth_in = tuple(shape_coords_in[a] for a in range(-len(shape_coords_in), 0) if a not in axis_input)
    orth_out = tuple(shape_coords_out[a] for a in range(-len(shape_coords_out), 0) if a not in axis_output)
    shape_orthogonal = np.broadcast_shapes(orth_in, orth_out)

    shape_in = _embed(shape_orthogonal, axis_input, {a: shape_coords_in[a] for a in axis_input})
    shape_out = _embed(shape_orthogonal, axis_output, {a: shape_coords_out[a] for a in axis_output})

    coordinates_input = tuple(np.broadcast_to(c, shape_in) for c in coordinates_input)
    coordinates_output = tuple(np.broadcast_to(c, shape_out) for c in coordinates_output)

    if perturb:
        # one generator for every coordinate, in order: x for ALL orthogonal slices, then y
        rng = np.random.default_rng(seed)
        jittered = []
        for c in coordinates_output:
            spread = np.ptp(c, axis=axis_output, keepdims=True)
            jittered.append(rng.normal(c, spread * EPSILON_PERTURB))
        coordinates_output = tuple(jittered)

    return (coordinates_input, coordinates_output, axis_input, axis_output,
            shape_coords_in, shape_coords_out, shape_orthogonal)
