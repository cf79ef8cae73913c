"""
Deterministic synthetic inputs shared by the golden-vector generator
(``tests/golden/make_golden.py``), the oracle tests and the GPU parity tests.

Grids used for GOLDEN comparisons are built from IEEE-basic operations only
(+, -, *, / on ``linspace`` output, with the rotation given by literal constants)
so that they are bit-identical on every host; the golden files carry a SHA-256 of
the input coordinates to prove it.  The benchmark family of
``benchmarks/regrid.py:15-25`` (which calls libm ``cos``/``sin``) is provided too,
for parity tests that compare the CUDA path with the oracle on the same process's
inputs (no cross-host determinism needed there).
"""

from __future__ import annotations

import hashlib
import pathlib

import numpy as np

ROOT_GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"

# cos(0.4), sin(0.4) as literals: no libm in the golden input path.
COS04 = 0.9210609940028851
SIN04 = 0.3894183423086505


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def curvilinear(nx: int, ny: int | None = None, distort: float = 0.0, c: float = COS04, s: float = SIN04,
                flip_x: bool = False, flip_y: bool = False, scale: float = 1.0, shift=(0.0, 0.0)):
    """Rotated + sheared (+ optionally distorted) curvilinear vertex grid, IEEE-basic ops only.

    Same family as ``benchmarks/regrid.py:15-25`` with the rotation angle 0.4 rad;
    ``distort`` adds a cubic wiggle that keeps the cells convex.
    """
    ny = nx if ny is None else ny
    x = np.linspace(-1, 1, nx)
    y = np.linspace(-1, 1, ny)
    if flip_x:
        x = x[::-1]
    if flip_y:
        y = y[::-1]
    x, y = np.meshgrid(x, y, indexing="ij")
    X = x * c - y * s + 0.05 * x * x
    Y = x * s + y * c + 0.05 * y * y
    if distort:
        X = X + distort * (4.0 * y * (1.0 - y * y))
        Y = Y + distort * (3.0 * x * (1.0 - x * x))
    X = X * scale + shift[0]
    Y = Y * scale + shift[1]
    return np.ascontiguousarray(X), np.ascontiguousarray(Y)


def rectilinear_over(X, Y, mx: int, my: int, shrink: float = 1.0):
    """Rectilinear output vertices spanning (a fraction of) the bounding box of (X, Y)."""
    x0, x1, y0, y1 = X.min(), X.max(), Y.min(), Y.max()
    if shrink != 1.0:
        cx, cy = 0.5 * (x0 + x1), 0.5 * (y0 + y1)
        hx, hy = 0.5 * (x1 - x0) * shrink, 0.5 * (y1 - y0) * shrink
        x0, x1, y0, y1 = cx - hx, cx + hx, cy - hy, cy + hy
    xo = np.linspace(x0, x1, mx)
    yo = np.linspace(y0, y1, my)
    xo, yo = np.meshgrid(xo, yo, indexing="ij")
    return np.ascontiguousarray(xo), np.ascontiguousarray(yo)


def benchmark_family(n: int, mx: int | None = None, my: int | None = None, distorted: bool = False,
                     angle: float = 0.4, phase: float = 0.0):
    """``benchmarks/regrid.py:15-25`` (libm cos/sin).  ``distorted`` adds the SURVEY §8d wiggle."""
    mx = n if mx is None else mx
    my = n if my is None else my
    x = np.linspace(-1, 1, n)
    y = np.linspace(-1, 1, n)
    x, y = np.meshgrid(x, y, indexing="ij")
    X = x * np.cos(angle) - y * np.sin(angle) + 0.05 * x * x
    Y = x * np.sin(angle) + y * np.cos(angle) + 0.05 * y * y
    if distorted:
        X = X + 0.01 * np.sin(3 * np.pi * y + phase)
        Y = Y + 0.01 * np.sin(2 * np.pi * x + phase)
    xo = np.linspace(X.min(), X.max(), mx)
    yo = np.linspace(Y.min(), Y.max(), my)
    xo, yo = np.meshgrid(xo, yo, indexing="ij")
    return (np.ascontiguousarray(X), np.ascontiguousarray(Y)), (np.ascontiguousarray(xo), np.ascontiguousarray(yo))


def perturb_like_reference(coords_output, axis_output, seed=42):
    """Restatement of the host-side jitter of ``regridding/_util.py:121-129`` for test inputs."""
    rng = np.random.default_rng(seed)
    out = []
    for coord in coords_output:
        ptp = np.ptp(coord, axis=axis_output, keepdims=True)
        out.append(rng.normal(coord, ptp * 1e-9))
    return tuple(out)


# --------------------------------------------------------------------------
# golden case table: name -> callable returning the kwargs of regridding.weights
# --------------------------------------------------------------------------

def _case_2d(name):
    if name == "fam40":
        gi = curvilinear(40)
        go = rectilinear_over(*gi, 48, 32)
    elif name == "fam100":  # BASELINE.json config 1 shape: 100x100 -> 120x80 vertices
        gi = curvilinear(100)
        go = rectilinear_over(*gi, 120, 80)
    elif name == "dist129":
        gi = curvilinear(129, distort=0.01)
        go = rectilinear_over(*gi, 129, 129)
    elif name == "dist257":
        gi = curvilinear(257, distort=0.01)
        go = rectilinear_over(*gi, 257, 257)
    elif name == "coarsen":
        gi = curvilinear(129, 97, distort=0.01)
        go = rectilinear_over(*gi, 33, 41)
    elif name == "refine":
        gi = curvilinear(33, 29, distort=0.01)
        go = rectilinear_over(*gi, 97, 101)
    elif name == "inner":  # output strictly inside the input grid
        gi = curvilinear(64, distort=0.01)
        go = rectilinear_over(*gi, 50, 60, shrink=0.45)
    elif name == "flipx":  # negative-orientation input grid
        gi = curvilinear(48, 40, flip_x=True)
        go = rectilinear_over(*gi, 40, 52)
    elif name == "curv2curv":  # both grids curvilinear, partial overlap
        gi = curvilinear(60, 50, distort=0.01)
        go = curvilinear(45, 55, c=0.9800665778412416, s=0.19866933079506122, scale=0.8, shift=(0.1, -0.05))
    elif name == "winput":
        gi = curvilinear(40, 36)
        go = rectilinear_over(*gi, 30, 44)
    else:
        raise KeyError(name)
    return gi, go


CASES_2D = ["fam40", "fam100", "dist129", "dist257", "coarsen", "refine", "inner", "flipx", "curv2curv", "winput"]
CASES_2D_FULL = ["fam40", "fam100", "coarsen", "refine", "inner", "flipx", "curv2curv", "winput"]  # arrays stored in full


def case_2d(name):
    gi, go = _case_2d(name)
    w = None
    if name == "winput":
        w = np.random.default_rng(7).random((gi[0].shape[0] - 1, gi[0].shape[1] - 1)) + 0.5
    return gi, go, w


def case_2d_batched():
    """Three frames, each with its own input AND output grid (orthogonal axis 0)."""
    gis, gos = [], []
    for f in range(3):
        c, s = [(COS04, SIN04), (0.9800665778412416, 0.19866933079506122),
                (0.9950041652780258, 0.09983341664682815)][f]
        gi = curvilinear(24, 20, distort=0.005 * f, c=c, s=s)
        go = rectilinear_over(*gi, 22, 26)
        gis.append(gi)
        gos.append(go)
    xi = np.stack([g[0] for g in gis])
    yi = np.stack([g[1] for g in gis])
    xo = np.stack([g[0] for g in gos])
    yo = np.stack([g[1] for g in gos])
    return (xi, yi), (xo, yo)


def cases_1d():
    """name -> (x_input (S, n), x_output (S, m), weights_input or None)."""
    rng = np.random.default_rng(11)
    out = {}
    base = np.linspace(4000.0, 7000.0, 513)
    S = 50
    xin = base * (1 + 1e-4 * rng.standard_normal((S, 1))) + 0.3 * ((base / 500.0) % 1.0) * rng.random((S, 1))
    xout = np.linspace(4001.0, 6999.0, 513) + 0.05 * rng.random((S, 1))
    out["spectra"] = (xin, xout, None)
    out["spectra_w"] = (xin[:5], xout[:5], rng.random((5, 512)) + 0.5)
    out["descending_uniform"] = (np.linspace(1, -1, 11)[None], np.linspace(-1.000001, 0.999999, 7)[None], None)
    out["descending_both"] = (np.linspace(1, -1, 11)[None], np.linspace(1.2, -0.7, 9)[None], None)
    out["descending_nonuniform"] = (np.array([[10.0, 6.0, 3.0, 1.0, 0.0]]), np.array([[0.0, 10.0]]), None)
    out["disjoint"] = (np.linspace(0, 1, 9)[None], np.linspace(2, 3, 5)[None], None)
    out["partial_left"] = (np.linspace(0, 1, 9)[None], np.linspace(0.33, 1.7, 6)[None], None)
    out["partial_right"] = (np.linspace(0, 1, 9)[None], np.linspace(-0.71, 0.52, 6)[None], None)
    out["coincident"] = (np.linspace(0, 1, 9)[None], np.linspace(0, 1, 5)[None], None)
    out["out_inside_in"] = (np.linspace(0, 1, 6)[None], np.linspace(0.21, 0.83, 14)[None], None)
    return out


def cases_multilinear_1d():
    """name -> (x_input, x_output, weights_input or None, bounds)."""
    rng = np.random.default_rng(21)
    out = {}
    x = np.linspace(0.0, 1.0, 11)
    out["single_sorted"] = (x, np.linspace(0.05, 0.95, 7), None, "extrapolate")
    out["single_unsorted_outside_extrapolate"] = (x, np.array([0.31, -0.2, 1.3, 0.0, 1.0, 0.5, 0.31, 0.77]), None, "extrapolate")
    out["single_unsorted_outside_nan"] = (x, np.array([0.31, -0.2, 1.3, 0.0, 1.0, 0.5, 0.31, 0.77]), None, "nan")
    base = np.linspace(4000.0, 7000.0, 65)
    S = 6
    xin = base * (1 + 1e-4 * rng.standard_normal((S, 1))) + 0.3 * ((base / 500.0) % 1.0) * rng.random((S, 1))
    xout = np.linspace(3990.0, 7010.0, 90) + 0.05 * rng.random((S, 1))
    out["spectra_stack_nan"] = (xin, xout, None, "nan")
    out["spectra_stack_extrapolate_w"] = (xin, rng.permuted(xout, axis=1), rng.random((S, 65)) + 0.5, "extrapolate")
    out["nonuniform"] = (np.cumsum(rng.random(40) + 0.1), np.sort(rng.random(55) * 25.0), None, "extrapolate")
    return out


def cases_interp_ndarray():
    """name -> (a, indices, kwargs) for ndarray_linear_interpolation (shapes of _tests/test_interp_ndarray.py:12-200)."""
    rng = np.random.default_rng(31)
    sz_t, sz_x, sz_y = 9, 10, 11
    out = {}
    a1 = rng.random(sz_x)
    out["1d_inside"] = (a1, (np.linspace(0, sz_x - 1, 21),), {})
    out["1d_extrapolate"] = (a1, (rng.random(40) * (sz_x + 3) - 2,), {"axis": -1})
    a2 = rng.random((sz_x, sz_y))
    x = np.linspace(0, sz_x - 1, 100)[:, None]
    y = np.linspace(0, sz_y - 1, 5)[None, :]
    out["2d_grid"] = (a2, tuple(np.broadcast_arrays(x, y)), {"axis": (0, ~0)})
    out["2d_scattered_extrapolate"] = (a2, (rng.random(500) * (sz_x + 2) - 1.5, rng.random(500) * (sz_y + 2) - 1.5), {})
    a3 = rng.random((sz_t, sz_x, sz_y))
    out["3d_axis12_shared_indices"] = (a3, (rng.random((1, 30)) * (sz_x - 1), rng.random((1, 30)) * (sz_y - 1)),
                                       {"axis": (1, 2), "axis_indices": (1,)})
    out["3d_axis12_per_slice_indices"] = (a3, (rng.random((sz_t, 7, 4)) * (sz_x - 1), rng.random((sz_t, 7, 4)) * (sz_y - 1)),
                                          {"axis": (1, 2), "axis_indices": (1, 2)})
    out["3d_axis0_1d"] = (a3, (rng.random((13, sz_x, sz_y)) * (sz_t - 1),), {"axis": 0, "axis_indices": 0})
    out["big_2d"] = (rng.random((200, 300)), (rng.random(20000) * 205 - 3, rng.random(20000) * 305 - 3), {})
    return out


def cases_find_indices():
    rng = np.random.default_rng(5)
    D, n, m = 7, 33, 41
    xin = np.sort(rng.random((D, n)), axis=1)
    xout = rng.random((D, m)) * 1.2 - 0.1
    xout[:, 0] = xin[:, 0]
    xout[:, 1] = xin[:, -1]
    xout[:, 2] = xin[:, 5]
    return xin, xout
