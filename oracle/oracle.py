"""
ctypes front-end of the CPU oracle (``oracle_regrid.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs.  The product package ``regridding_b200`` never
imports this module.

The functions mirror the *kernel call sites* of the reference:
``weights_conservative_2d`` (c2d.py:80-126), ``_coalesce`` (warr.py:44-73),
``_regrid_from_weights`` (rfw.py:165-182), ``_weights_conservative_1d``
(c1d.py:60-189) and the 1D ``find_indices`` kernels.
"""

from __future__ import annotations

import ctypes
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle_regrid.so"

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> pathlib.Path:
    """Compile the oracle with the recipe in ``oracle/Makefile``."""
    src = _HERE / "oracle_regrid.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = ctypes.CDLL(str(_LIB_PATH))
        L.orc_set_mode.argtypes = [ctypes.c_int]
        L.orc_get_mode.restype = ctypes.c_int
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_point_in_polygon.argtypes = [ctypes.c_double, ctypes.c_double, _c_double_p, _c_double_p, ctypes.c_int64]
        L.orc_point_in_polygon.restype = ctypes.c_int
        L.orc_grid_volume.argtypes = [_c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_int64, _c_double_p]
        for f in (L.orc_index_of_point_secant, L.orc_index_of_point_brute):
            f.argtypes = [_c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_int64,
                          ctypes.c_double, ctypes.c_double, _c_int64_p]
        L.orc_weights_conservative_2d.argtypes = [
            _c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_int64,
            _c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_int64,
            _c_double_p,
            ctypes.POINTER(_c_int64_p), ctypes.POINTER(_c_int64_p), ctypes.POINTER(_c_double_p),
        ]
        L.orc_weights_conservative_2d.restype = ctypes.c_int64
        L.orc_coalesce.argtypes = [ctypes.c_int64, _c_int64_p, _c_int64_p, _c_double_p,
                                   _c_int64_p, _c_int64_p, _c_double_p]
        L.orc_coalesce.restype = ctypes.c_int64
        L.orc_reduceat_segment.argtypes = [_c_double_p, ctypes.c_int64]
        L.orc_reduceat_segment.restype = ctypes.c_double
        L.orc_regrid_from_weights.argtypes = [ctypes.c_int64, _c_int64_p, _c_int64_p, _c_double_p,
                                              ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                              _c_double_p, _c_double_p]
        L.orc_weights_conservative_1d.argtypes = [_c_double_p, ctypes.c_int64, _c_double_p, ctypes.c_int64,
                                                  _c_double_p, _c_int64_p, _c_int64_p, _c_double_p]
        L.orc_weights_conservative_1d.restype = ctypes.c_int64
        L.orc_weights_conservative_1d_batched.argtypes = [
            ctypes.c_int64, _c_double_p, ctypes.c_int64, _c_double_p, ctypes.c_int64, _c_double_p,
            _c_int64_p, _c_int64_p, _c_double_p, _c_int64_p]
        for f in (L.orc_find_indices_brute_1d, L.orc_find_indices_searchsorted_1d):
            f.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _c_double_p, _c_double_p,
                          ctypes.c_int64, _c_int64_p]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_c_double_p)


def _i(a):
    return a.ctypes.data_as(_c_int64_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_mode(mode: str | int) -> None:
    """``"jit"`` (default; Numba fastmath contraction pattern) or ``"strict"`` (plain IEEE)."""
    if isinstance(mode, str):
        mode = {"jit": 1, "strict": 0}[mode]
    lib().orc_set_mode(int(mode))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int | None = None) -> int:
    """Use ``n`` OpenMP threads (default: every core this process may run on), whatever ``OMP_NUM_THREADS``
    says -- ``torch.distributed.run`` exports ``OMP_NUM_THREADS=1`` to its workers."""
    import os

    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    L = lib()
    L.orc_set_num_threads.argtypes = [ctypes.c_int]
    L.orc_set_num_threads(int(n))
    return num_threads()


def index_of_points(x, y, px, py, fill_value: int, method: str = "secant") -> np.ndarray:
    """Flat cell index ``i * (ny - 1) + j`` of the lowest-index cell containing each point, else ``fill_value``
    (c2d/_grids.py:223-279 brute, 356-463 secant), OpenMP over the points."""
    x, y = _f64(x), _f64(y)
    px, py = _f64(px).reshape(-1), _f64(py).reshape(-1)
    out = np.empty(px.size, dtype=np.int64)
    L = lib()
    L.orc_index_of_points.argtypes = [ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int64, _c_double_p, _c_double_p, ctypes.c_int64, _c_int64_p]
    L.orc_index_of_points.restype = None
    L.orc_index_of_points(0 if method == "secant" else 1, _d(x), _d(y), x.shape[0], x.shape[1], px.size,
                          _d(px), _d(py), int(fill_value), _i(out))
    return out


def point_is_inside_polygon(x, y, vertices_x, vertices_y) -> bool:
    vx, vy = _f64(vertices_x), _f64(vertices_y)
    return bool(lib().orc_point_in_polygon(float(x), float(y), _d(vx), _d(vy), vx.size))


def grid_volume(x, y) -> np.ndarray:
    x, y = _f64(x), _f64(y)
    nx, ny = x.shape
    out = np.empty((nx - 1, ny - 1))
    lib().orc_grid_volume(_d(x), _d(y), nx, ny, _d(out))
    return out


def index_of_point(x, y, px, py, method="secant") -> tuple[int, int]:
    x, y = _f64(x), _f64(y)
    out = np.empty(2, dtype=np.int64)
    f = lib().orc_index_of_point_secant if method == "secant" else lib().orc_index_of_point_brute
    f(_d(x), _d(y), x.shape[0], x.shape[1], float(px), float(py), _i(out))
    return int(out[0]), int(out[1])


def weights_conservative_2d(grid_input, grid_output, weights_input=None):
    """Raw (uncoalesced) triplets in the reference's emission order (c2d.py:80-126)."""
    xi, yi = (_f64(a) for a in grid_input)
    xo, yo = (_f64(a) for a in grid_output)
    w = None if weights_input is None else _f64(weights_input)
    pii, pio, pv = _c_int64_p(), _c_int64_p(), _c_double_p()
    n = lib().orc_weights_conservative_2d(
        _d(xi), _d(yi), xi.shape[0], xi.shape[1],
        _d(xo), _d(yo), xo.shape[0], xo.shape[1],
        _d(w) if w is not None else None,
        ctypes.byref(pii), ctypes.byref(pio), ctypes.byref(pv),
    )
    try:
        ii = np.ctypeslib.as_array(pii, shape=(max(n, 1),))[:n].copy()
        io = np.ctypeslib.as_array(pio, shape=(max(n, 1),))[:n].copy()
        v = np.ctypeslib.as_array(pv, shape=(max(n, 1),))[:n].copy()
    finally:
        lib().orc_free(pii)
        lib().orc_free(pio)
        lib().orc_free(pv)
    return ii, io, v


def coalesce(indices_input, indices_output, values):
    """warr.py:44-73."""
    ii = np.ascontiguousarray(indices_input, dtype=np.int64)
    io = np.ascontiguousarray(indices_output, dtype=np.int64)
    v = _f64(values)
    n = v.size
    oi, oo, ov = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float64)
    m = lib().orc_coalesce(n, _i(ii), _i(io), _d(v), _i(oi), _i(oo), _d(ov))
    return oi[:m].copy(), oo[:m].copy(), ov[:m].copy()


def reduceat_segment(a) -> float:
    a = _f64(a)
    return float(lib().orc_reduceat_segment(_d(a), a.size))


def regrid_from_weights(indices_input, indices_output, values, values_input, n_out: int) -> np.ndarray:
    """rfw.py:165-182 with one set of weights shared by all D slices; ``values_input`` is (D, n_in)."""
    ii = np.ascontiguousarray(indices_input, dtype=np.int64)
    io = np.ascontiguousarray(indices_output, dtype=np.int64)
    v = _f64(values)
    vin = _f64(values_input)
    D, n_in = vin.shape
    out = np.zeros((D, n_out))
    lib().orc_regrid_from_weights(v.size, _i(ii), _i(io), _d(v), D, n_in, n_out, _d(vin), _d(out))
    return out


def weights_conservative_1d(x_input, x_output, weights_input=None):
    """c1d.py:60-189 for one spectrum."""
    xi, xo = _f64(x_input), _f64(x_output)
    w = None if weights_input is None else _f64(weights_input)
    cap = xi.size + xo.size
    ii, io, v = np.empty(cap, np.int64), np.empty(cap, np.int64), np.empty(cap, np.float64)
    n = lib().orc_weights_conservative_1d(_d(xi), xi.size, _d(xo), xo.size,
                                          _d(w) if w is not None else None, _i(ii), _i(io), _d(v))
    return ii[:n].copy(), io[:n].copy(), v[:n].copy()


def weights_conservative_1d_batched(x_input, x_output, weights_input=None):
    """wcons.py:59-106: (S, n) and (S, m) stacks -> list of S triplet tuples."""
    xi, xo = _f64(x_input), _f64(x_output)
    S, n = xi.shape
    m = xo.shape[1]
    w = None if weights_input is None else _f64(weights_input)
    cap = n + m
    ii, io = np.empty((S, cap), np.int64), np.empty((S, cap), np.int64)
    v = np.empty((S, cap), np.float64)
    counts = np.empty(S, np.int64)
    lib().orc_weights_conservative_1d_batched(S, _d(xi), n, _d(xo), m, _d(w) if w is not None else None,
                                              _i(ii), _i(io), _d(v), _i(counts))
    return [(ii[s, :c].copy(), io[s, :c].copy(), v[s, :c].copy()) for s, c in enumerate(counts)]


def find_indices_1d(x_input, x_output, fill_value: int, method: str = "brute") -> np.ndarray:
    """fib.py:24-51 / fis.py:24-62 on (D, n) and (D, m) stacks."""
    xi, xo = _f64(x_input), _f64(x_output)
    D, n = xi.shape
    m = xo.shape[1]
    out = np.empty((D, m), np.int64)
    f = lib().orc_find_indices_brute_1d if method == "brute" else lib().orc_find_indices_searchsorted_1d
    f(D, n, m, _d(xi), _d(xo), int(fill_value), _i(out))
    return out


def transpose_weights_conservative_values(indices_input, indices_output, values, volume_input, volume_output,
                                          weights_input=None) -> np.ndarray:
    """wT.py:236-249 for one element: NumPy's own evaluation order (the reference IS NumPy here);
    volumes from ``grid_volume`` (2D) or ``np.diff`` (= cell_length, c1d/_grids.py:10-35)."""
    v = np.array(values, dtype=np.float64)
    if weights_input is not None:
        v = v / np.square(weights_input[indices_input])
    return v * volume_input[indices_input] / volume_output[indices_output]


def multilinear2d_weights(x, y, px, py, cells_flat):
    """NumPy restatement of the 2D multilinear (bilinear) rule (test infrastructure; the reference has no 2D
    multilinear, regridding/_weights/_weights_multilinear.py:128-131 raises -- this extends its 1D rule,
    wml.py:105-119, 185-202): for every point, (u, v) of the bilinear map of its containing cell (flat cell
    index ``cells_flat``, u along axis 0) by Newton from (1/2, 1/2), and the weights
    ``(1-u)(1-v), (1-u)v, u(1-v), uv`` at the flat vertex indices ``a, a+1, a+ny, a+ny+1`` (ascending).
    Returns ``(idx4 int64 [P, 4], w4 float64 [P, 4])``.  Points with a negative cell index get NaN weights."""
    x, y = _f64(x), _f64(y)
    px, py = _f64(px).reshape(-1), _f64(py).reshape(-1)
    cells = np.asarray(cells_flat, dtype=np.int64).reshape(-1)
    nx, ny = x.shape
    ncy = ny - 1
    ok = cells >= 0
    c = np.where(ok, cells, 0)
    ci, cj = c // ncy, c % ncy
    a = ci * ny + cj
    xf, yf = x.reshape(-1), y.reshape(-1)
    x00, x01, x10, x11 = xf[a], xf[a + 1], xf[a + ny], xf[a + ny + 1]
    y00, y01, y10, y11 = yf[a], yf[a + 1], yf[a + ny], yf[a + ny + 1]
    ax, bx, cx = x10 - x00, x01 - x00, ((x00 - x10) - x01) + x11
    ay, by, cy = y10 - y00, y01 - y00, ((y00 - y10) - y01) + y11
    u = np.full(px.shape, 0.5)
    v = np.full(px.shape, 0.5)
    active = np.ones(px.shape, dtype=bool)
    prev = np.full(px.shape, np.inf)
    with np.errstate(all="ignore"):
        for _ in range(24):
            ex = (((x00 + u * ax) + v * bx) + (u * v) * cx) - px
            ey = (((y00 + u * ay) + v * by) + (u * v) * cy) - py
            xu, xv = ax + v * cx, bx + u * cx
            yu, yv = ay + v * cy, by + u * cy
            det = xu * yv - xv * yu
            du = (yv * ex - xv * ey) / det
            dv = (xu * ey - yu * ex) / det
            u = np.where(active, u - du, u)
            v = np.where(active, v - dv, v)
            # converged, or stagnating at the rounding noise of the residual
            step = np.fmax(np.abs(du), np.abs(dv))
            stop = (step < 1e-14) | ((step < 1e-6) & (step >= prev))
            prev = np.where(active, step, prev)
            active &= ~stop
            if not active.any():
                break
    w4 = np.stack(((1 - u) * (1 - v), (1 - u) * v, u * (1 - v), u * v), axis=1)
    w4[~ok] = np.nan
    idx4 = np.stack((a, a + 1, a + ny, a + ny + 1), axis=1)
    return idx4, w4


def fill_gauss_seidel_2d(a, where, num_iterations: int) -> np.ndarray:
    """``_fill_gauss_seidel_2d`` (regridding/_fill/_gauss_seidel.py:83-139): ``a`` (T, ny, nx) with the guess
    already in place, ``where`` boolean (T, ny, nx); returns the relaxed copy."""
    L = lib()
    a = np.array(a, dtype=np.float64, order="C", copy=True)
    w = np.ascontiguousarray(where, dtype=np.uint8)
    assert a.ndim == 3 and w.shape == a.shape
    L.orc_fill_gauss_seidel_2d.argtypes = [_c_double_p, ctypes.POINTER(ctypes.c_ubyte), ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_int64, ctypes.c_int64]
    L.orc_fill_gauss_seidel_2d.restype = None
    L.orc_fill_gauss_seidel_2d(_d(a), w.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), a.shape[0], a.shape[1],
                               a.shape[2], int(num_iterations))
    return a


def fill(a, where=None, axis=None, guess=None, num_iterations: int = 100) -> np.ndarray:
    """``regridding.fill(method="gauss_seidel")`` host logic (``_fill.py:10-109``, ``_gauss_seidel.py:13-81``)
    around the C relaxation: NaN mask by default, median guess, axis bookkeeping."""
    a = np.array(a, dtype=np.float64, copy=True)
    if where is None:
        where = np.isnan(a)
    a, where = np.broadcast_arrays(a, where)
    a = a.copy()
    nd = a.ndim
    ax = tuple(range(nd)) if axis is None else tuple(np.lib.array_utils.normalize_axis_tuple(axis, nd))
    if len(ax) != 2:
        raise ValueError(f"The number of interpolation axes, {len(ax)},is not supported")
    if guess is None:
        masked = np.where(where, np.nan, a)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            guess = np.nanmedian(masked, axis=ax, keepdims=True)
        guess = np.where(np.isnan(guess), 0, guess)
    a[where] = np.broadcast_to(guess, a.shape)[where]
    am = np.moveaxis(a, ax, (-2, -1))
    wm = np.moveaxis(where, ax, (-2, -1))
    shape_moved = am.shape
    res = fill_gauss_seidel_2d(am.reshape(-1, *shape_moved[-2:]), wm.reshape(-1, *shape_moved[-2:]), num_iterations)
    return np.moveaxis(res.reshape(shape_moved), (-2, -1), ax)
