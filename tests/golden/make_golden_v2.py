"""
Golden vectors of ``tests/golden/golden_v2.npz`` (transposed weights), made by RUNNING THE
REFERENCE ITSELF (sun-data/regridding through Numba) in the build container:

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden_v2.py

Same rules as make_golden.py: inputs from ``tests/cases.py``; the ``.npz`` is committed and
travels to the GPU box, ``/root/reference`` does not.
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


def transposed_2d():
    for name in ("fam40", "winput", "coarsen"):
        gi, go, w = cases.case_2d(name)
        W = regridding.weights(gi, go, weights_input=w, method="conservative")
        vals = np.random.default_rng(0).random((3, *W[1]))
        fwd = regridding.regrid_from_weights(*W, vals)
        # plain transpose (wT.py:13-52)
        Wt = regridding.transpose_weights(W)
        put(f"t2d/{name}/plain_sha", cases.sha(*Wt[0][()]))
        put(f"t2d/{name}/plain_apply", regridding.regrid_from_weights(*Wt, fwd))
        # conservative transpose (wT.py:55-262)
        Wc = regridding.transpose_weights_conservative(W, gi, go, weights_input=w)
        ii, io, v = Wc[0][()]
        put(f"t2d/{name}/cons_ii", ii.astype(np.int32))
        put(f"t2d/{name}/cons_io", io.astype(np.int32))
        put(f"t2d/{name}/cons_v", v)
        put(f"t2d/{name}/cons_shapes", np.array([*Wc[1], *Wc[2]]))
        put(f"t2d/{name}/cons_apply", regridding.regrid_from_weights(*Wc, fwd))


def transposed_batched():
    gi, go = cases.case_2d_batched()
    kw = dict(axis_input=(1, 2), axis_output=(1, 2))
    W = regridding.weights(gi, go, method="conservative", **kw)
    Wc = regridding.transpose_weights_conservative(W, gi, go, **kw)
    for f in range(3):
        put(f"t2d_batched/{f}/v", Wc[0][f][2])
        put(f"t2d_batched/{f}/sha", cases.sha(*Wc[0][f]))
    vals = np.random.default_rng(0).random(W[1])
    fwd = regridding.regrid_from_weights(*W, vals, **kw)
    put("t2d_batched/apply", regridding.regrid_from_weights(*Wc, fwd, **kw))


def transposed_1d():
    for name in ("spectra", "spectra_w", "descending_nonuniform", "descending_both"):
        if name not in cases.cases_1d():
            continue
        xin, xout, w = cases.cases_1d()[name]
        W = regridding.weights((xin,), (xout,), axis_input=-1, axis_output=-1, weights_input=w, method="conservative")
        Wc = regridding.transpose_weights_conservative(W, (xin,), (xout,), axis_input=-1, axis_output=-1,
                                                       weights_input=w)
        flat = Wc[0].reshape(-1)
        put(f"t1d/{name}/v", np.concatenate([e[2] for e in flat]))
        put(f"t1d/{name}/ii", np.concatenate([e[0] for e in flat]))
        vals = np.random.default_rng(0).random(W[1])
        fwd = regridding.regrid_from_weights(*W, vals, axis_input=-1, axis_output=-1)
        put(f"t1d/{name}/apply", regridding.regrid_from_weights(*Wc, fwd, axis_input=-1, axis_output=-1))


if __name__ == "__main__":
    transposed_2d()
    transposed_batched()
    transposed_1d()
    out = HERE / "golden_v2.npz"
    np.savez_compressed(out, **G)
    print(f"wrote {out}: {len(G)} arrays, {out.stat().st_size / 1e6:.2f} MB")
