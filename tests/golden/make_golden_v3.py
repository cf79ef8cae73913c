"""
Golden vectors of ``tests/golden/golden_v3.npz`` (1D multilinear weights -- ``weights()``'s default method --
and their application), made by RUNNING THE REFERENCE ITSELF (sun-data/regridding through Numba) in the build
container:

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden_v3.py

Same rules as make_golden.py: inputs from ``tests/cases.py``; the ``.npz`` is committed and travels to the GPU
box, ``/root/reference`` does not.
"""

from __future__ import annotations

import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")

import regridding  # noqa: E402  (the reference)

from tests import cases  # noqa: E402

G: dict[str, np.ndarray] = {}


def put(key, value):
    G[key] = np.asarray(value)


def multilinear_1d():
    for name, (x_in, x_out, w, bounds) in cases.cases_multilinear_1d().items():
        kw = dict(axis_input=-1, axis_output=-1) if x_in.ndim > 1 else {}
        W = regridding.weights((x_in,), (x_out,), weights_input=w, method="multilinear", bounds=bounds, **kw)
        flat = W[0].reshape(-1)
        for d in range(flat.size):
            ii, io, v = flat[d]
            put(f"ml1d/{name}/{d}/ii", ii.astype(np.int32))
            put(f"ml1d/{name}/{d}/io", io.astype(np.int32))
            put(f"ml1d/{name}/{d}/v", v)
        put(f"ml1d/{name}/shapes", np.array([*W[1], *W[2]]))
        vals = np.random.default_rng(3).random(W[1])
        put(f"ml1d/{name}/apply", regridding.regrid_from_weights(*W, vals, **kw))
        put(f"ml1d/{name}/regrid", regridding.regrid((x_in,), (x_out,), vals, method="multilinear", bounds=bounds, **kw))


def interp_ndarray():
    for name, (a, indices, kw) in cases.cases_interp_ndarray().items():
        put(f"interp/{name}", regridding.ndarray_linear_interpolation(a, indices, **kw))


if __name__ == "__main__":
    multilinear_1d()
    interp_ndarray()
    np.savez_compressed(HERE / "golden_v3.npz", **G)
    print("wrote", len(G), "arrays")
