"""
Generate the golden vectors of ``tests/golden/golden_v1.npz`` by RUNNING THE
REFERENCE ITSELF (sun-data/regridding through Numba) in the build container.

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden.py

Needs ``/root/reference`` (read-only mount) and numba; it therefore runs only in
the build container -- the resulting ``.npz`` is committed and is what travels to
the GPU box.  Inputs come from ``tests/cases.py`` (IEEE-basic operations only) and
their SHA-256 is stored beside every result.
"""

from __future__ import annotations

import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")

import regridding  # noqa: E402  (the reference)
from regridding import _util  # noqa: E402
from regridding._weights._weights_conservative_2d import weights_conservative_2d  # noqa: E402

from tests import cases  # noqa: E402

G: dict[str, np.ndarray] = {}


def put(key, value):
    G[key] = np.asarray(value)


def golden_2d():
    for name in cases.CASES_2D:
        gi, go, w = cases.case_2d(name)
        put(f"c2d/{name}/input_sha", cases.sha(*gi, *go))
        weights, shape_in, shape_out = regridding.weights(gi, go, weights_input=w, method="conservative")
        ii, io, v = weights[()]
        put(f"c2d/{name}/shape_in", shape_in)
        put(f"c2d/{name}/shape_out", shape_out)
        put(f"c2d/{name}/nnz", v.size)
        put(f"c2d/{name}/final_sha", cases.sha(ii, io, v))
        put(f"c2d/{name}/final_index_sha", cases.sha(ii, io))
        put(f"c2d/{name}/sum_v", v.sum())
        if name in cases.CASES_2D_FULL:
            put(f"c2d/{name}/ii", ii.astype(np.int32))
            put(f"c2d/{name}/io", io.astype(np.int32))
            put(f"c2d/{name}/v", v)
        else:
            put(f"c2d/{name}/v_sample", v[::61])
            put(f"c2d/{name}/ii_sample", ii[::61])
            put(f"c2d/{name}/io_sample", io[::61])
        # raw emission-order triplets of the kernel call site (c2d.py:80-126)
        ci, co, *_ = _util._normalize_input_output_coordinates(gi, go, perturb=True, seed=42)
        ci = tuple(np.ascontiguousarray(a) for a in ci)
        co = tuple(np.ascontiguousarray(a) for a in co)
        put(f"c2d/{name}/perturbed_sha", cases.sha(*co))
        rii, rio, rv = weights_conservative_2d(ci, co, w)
        put(f"c2d/{name}/raw_n", rv.size)
        put(f"c2d/{name}/raw_sha", cases.sha(rii, rio, rv))
        put(f"c2d/{name}/raw_index_sha", cases.sha(rii, rio))
        # apply: three frames sharing the weights
        vals = np.random.default_rng(0).random((3, *shape_in))
        res = regridding.regrid_from_weights(weights, shape_in, shape_out, vals)
        put(f"c2d/{name}/apply_sha", cases.sha(res))
        put(f"c2d/{name}/apply_sum", res.sum())
        if name in ("fam40", "winput"):
            put(f"c2d/{name}/apply", res)
        print(name, "nnz", v.size, "raw", rv.size, flush=True)


def golden_2d_batched():
    gi, go = cases.case_2d_batched()
    put("c2d_batched/input_sha", cases.sha(*gi, *go))
    weights, shape_in, shape_out = regridding.weights(gi, go, axis_input=(1, 2), axis_output=(1, 2),
                                                      method="conservative")
    put("c2d_batched/shape_in", shape_in)
    put("c2d_batched/shape_out", shape_out)
    for f in range(3):
        ii, io, v = weights[f]
        put(f"c2d_batched/{f}/ii", ii.astype(np.int32))
        put(f"c2d_batched/{f}/io", io.astype(np.int32))
        put(f"c2d_batched/{f}/v", v)
    vals = np.random.default_rng(0).random(shape_in)
    res = regridding.regrid_from_weights(weights, shape_in, shape_out, vals, axis_input=(1, 2), axis_output=(1, 2))
    put("c2d_batched/apply", res)
    # seeds (regridding/_weights/_weights_test.py:39-103)
    w7, *_ = regridding.weights(gi, go, axis_input=(1, 2), axis_output=(1, 2), method="conservative", seed=7)
    put("c2d_batched/seed7_v0", w7[0][2])
    wnp, *_ = regridding.weights(gi, go, axis_input=(1, 2), axis_output=(1, 2), method="conservative", perturb=False)
    put("c2d_batched/noperturb_v0", wnp[0][2])
    put("c2d_batched/noperturb_ii0", wnp[0][0].astype(np.int32))
    put("c2d_batched/noperturb_io0", wnp[0][1].astype(np.int32))


def golden_1d():
    for name, (xin, xout, w) in cases.cases_1d().items():
        put(f"c1d/{name}/input_sha", cases.sha(xin, xout))
        weights, shape_in, shape_out = regridding.weights((xin,), (xout,), axis_input=-1, axis_output=-1,
                                                          weights_input=w, method="conservative")
        flat = weights.reshape(-1)
        put(f"c1d/{name}/counts", [e[2].size for e in flat])
        put(f"c1d/{name}/ii", np.concatenate([e[0] for e in flat]).astype(np.int64))
        put(f"c1d/{name}/io", np.concatenate([e[1] for e in flat]).astype(np.int64))
        put(f"c1d/{name}/v", np.concatenate([e[2] for e in flat]))
        vals = np.random.default_rng(0).random(shape_in)
        res = regridding.regrid_from_weights(weights, shape_in, shape_out, vals, axis_input=-1, axis_output=-1)
        put(f"c1d/{name}/apply", res)
        print("1d", name, [e[2].size for e in flat][:3], flush=True)


def golden_find_indices():
    xin, xout = cases.cases_find_indices()
    put("find/input_sha", cases.sha(xin, xout))
    for method in ("brute", "searchsorted"):
        (r,) = regridding.find_indices((xin,), (xout,), axis_input=-1, axis_output=-1, method=method)
        put(f"find/{method}", r)
        (r,) = regridding.find_indices((xin,), (xout,), axis_input=-1, axis_output=-1, method=method, fill_value=-1)
        put(f"find/{method}_fillm1", r)


def golden_primitives():
    """Truth tables of the geometric predicates on a fixed point set (geom.py, grids2.py)."""
    from regridding._weights._weights_conservative_2d import _grids
    import numba

    gi, _, _ = cases.case_2d("fam40")
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(gi[0].min(), gi[0].max(), 400), rng.uniform(gi[1].min(), gi[1].max(), 400)], axis=1)
    # include exact vertices and edge midpoints
    pts[:20, 0] = gi[0].reshape(-1)[::80][:20]
    pts[:20, 1] = gi[1].reshape(-1)[::80][:20]

    @numba.njit
    def locate(points, gx, gy, brute):
        out = np.empty((points.shape[0], 2), dtype=np.int64)
        for k in range(points.shape[0]):
            if brute:
                i, j = _grids.index_of_point_brute((points[k, 0], points[k, 1]), (gx, gy))
            else:
                i, j = _grids.index_of_point_secant((points[k, 0], points[k, 1]), (gx, gy))
            out[k, 0] = i
            out[k, 1] = j
        return out

    put("prim/points", pts)
    put("prim/locate_brute", locate(pts, gi[0], gi[1], True))
    put("prim/locate_secant", locate(pts, gi[0], gi[1], False))

    @numba.njit(fastmath=True)  # grid_volume is inlined into a fastmath=True caller (c2d.py:76-102)
    def vol(gx, gy):
        return _grids.grid_volume((gx, gy))

    for name in ("fam40", "coarsen", "flipx"):
        g, _, _ = cases.case_2d(name)
        put(f"prim/volume_{name}", vol(*g))

    # unit square, both orientations, incl. edges and vertices (regridding/_tests/test_geometry.py:432-503)
    sq_x = np.array([0.0, 1.0, 1.0, 0.0])
    sq_y = np.array([0.0, 0.0, 1.0, 1.0])
    q = np.array([[0.5, 0.5], [0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0], [1, 0.5], [0.5, 1], [0, 0.5],
                  [1.5, 0.5], [-0.5, 0.5], [0.5, 1.5], [0.5, -0.5], [2, 2], [1, 2], [0, -1]], dtype=float)

    @numba.njit
    def pip(points, vx, vy):
        out = np.empty(points.shape[0], dtype=np.bool_)
        for k in range(points.shape[0]):
            out[k] = regridding.geometry.point_is_inside_polygon(points[k, 0], points[k, 1], vx, vy)
        return out

    put("prim/pip_points", q)
    put("prim/pip_ccw", pip(q, sq_x, sq_y))
    put("prim/pip_cw", pip(q, sq_x[::-1].copy(), sq_y[::-1].copy()))


if __name__ == "__main__":
    golden_primitives()
    golden_find_indices()
    golden_1d()
    golden_2d_batched()
    golden_2d()
    out = HERE / "golden_v1.npz"
    np.savez_compressed(out, **G)
    print("wrote", out, out.stat().st_size / 1e6, "MB")
