"""2D multilinear (bilinear) regridding on curvilinear vertex grids -- BASELINE config 5.

The reference has no 2D multilinear (it raises, wml.py:128-131), so parity is pinned on (i) the NumPy
restatement in ``oracle/oracle.py`` (``multilinear2d_weights``), itself checked on the CPU against the
reference's 1D rule on rectilinear grids and against exactness for functions linear in (x, y), and
(ii) the same properties evaluated through the CUDA path at config-5 size."""

import numpy as np
import pytest

from tests import cases


def _rectilinear(nx, ny, seed=0):
    rng = np.random.default_rng(seed)
    gx = np.cumsum(rng.random(nx) + 0.5)
    gy = np.cumsum(rng.random(ny) + 0.5)
    return np.meshgrid(gx, gy, indexing="ij"), gx, gy


# ----------------------------------------------------------------------------------------------
# CPU: the oracle restatement
# ----------------------------------------------------------------------------------------------
def test_oracle_multilinear2d_is_separable_on_rectilinear_grids():
    """On a rectilinear grid the bilinear weights are the products of the reference's 1D weights
    w1 = (x - x0) / (x1 - x0), w0 = 1 - w1 (wml.py:185-186) along each axis."""
    from oracle import oracle

    (x, y), gx, gy = _rectilinear(9, 7)
    rng = np.random.default_rng(1)
    px = rng.uniform(gx[0], gx[-1], 200)
    py = rng.uniform(gy[0], gy[-1], 200)
    i = np.clip(np.searchsorted(gx, px) - 1, 0, gx.size - 2)
    j = np.clip(np.searchsorted(gy, py) - 1, 0, gy.size - 2)
    idx4, w4 = oracle.multilinear2d_weights(x, y, px, py, i * (gy.size - 1) + j)
    u = (px - gx[i]) / (gx[i + 1] - gx[i])
    v = (py - gy[j]) / (gy[j + 1] - gy[j])
    expect = np.stack(((1 - u) * (1 - v), (1 - u) * v, u * (1 - v), u * v), axis=1)
    assert np.allclose(w4, expect, rtol=0, atol=1e-13)
    a = i * gy.size + j
    assert np.array_equal(idx4, np.stack((a, a + 1, a + gy.size, a + gy.size + 1), axis=1))
    assert np.allclose(w4.sum(axis=1), 1.0, rtol=0, atol=1e-14)


def test_oracle_multilinear2d_reproduces_linear_functions_on_curvilinear_grids():
    from oracle import oracle

    x, y = cases.curvilinear(17, 13, distort=0.02)
    rng = np.random.default_rng(2)
    # points inside random cells: bilinear image of random (u, v)
    ci, cj = rng.integers(0, 16, 300), rng.integers(0, 12, 300)
    u, v = rng.random(300), rng.random(300)
    corner = lambda g, di, dj: g[ci + di, cj + dj]  # noqa: E731
    px = (corner(x, 0, 0) * (1 - u) + corner(x, 1, 0) * u) * (1 - v) + (corner(x, 0, 1) * (1 - u) + corner(x, 1, 1) * u) * v
    py = (corner(y, 0, 0) * (1 - u) + corner(y, 1, 0) * u) * (1 - v) + (corner(y, 0, 1) * (1 - u) + corner(y, 1, 1) * u) * v
    idx4, w4 = oracle.multilinear2d_weights(x, y, px, py, ci * 12 + cj)
    f = 0.3 + 1.7 * x - 0.9 * y
    got = (w4 * f.reshape(-1)[idx4]).sum(axis=1)
    assert np.allclose(got, 0.3 + 1.7 * px - 0.9 * py, rtol=0, atol=1e-12)
    assert np.allclose(w4[:, 2] + w4[:, 3], u, rtol=0, atol=1e-10)  # u along axis 0


# ----------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI
# ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def env():
    import torch

    import regridding_b200 as rg
    from oracle import oracle

    return rg, torch, torch.device("cuda", 0), oracle


@pytest.mark.gpu
def test_multilinear2d_weights_match_oracle(env):
    rg, torch, dev, oracle = env
    x, y = cases.curvilinear(33, 29, distort=0.01)
    go = cases.rectilinear_over(x, y, 41, 37, shrink=0.5)  # all inside
    px, py = go[0].reshape(-1), go[1].reshape(-1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    idx4, w4, n_out = rg.device.multilinear2d_weights(t(x), t(y), t(px), t(py), "nan")
    cells = rg.device.find_indices_2d(t(x), t(y), t(px), t(py), -1).cpu().numpy()
    assert int(n_out.item()) == 0 and (cells >= 0).all()
    oi, ow = oracle.multilinear2d_weights(x, y, px, py, cells)
    assert np.array_equal(idx4.cpu().numpy(), oi)
    assert np.allclose(w4.cpu().numpy(), ow, rtol=0, atol=1e-12)  # tolerance: 1e-12 absolute on weights in [0, 1]


@pytest.mark.gpu
def test_multilinear2d_bounds_modes(env):
    rg, torch, dev, oracle = env
    x, y = cases.curvilinear(21, 21)
    go = cases.rectilinear_over(x, y, 30, 30)  # bounding box of the rotated grid: corners fall outside
    f = 2.0 - 0.5 * x + 0.25 * y
    exact = 2.0 - 0.5 * go[0] + 0.25 * go[1]
    inside = np.array([[oracle.index_of_point(x, y, go[0][i, j], go[1][i, j], "brute")[0] < 10**9
                        for j in range(30)] for i in range(30)])
    assert 0 < inside.sum() < inside.size
    r_nan = rg.regrid((x, y), go, f, method="multilinear", bounds="nan")
    assert np.array_equal(np.isnan(r_nan), ~inside)
    assert np.allclose(r_nan[inside], exact[inside], rtol=0, atol=1e-12)
    # extrapolation continues the bilinear map of the nearest border cell: exact for a linear field on a
    # grid whose border cells are parallelograms up to the shear term -- check it is finite and close
    r_ext = rg.regrid((x, y), go, f, method="multilinear", bounds="extrapolate")
    assert np.isfinite(r_ext).all()
    assert np.allclose(r_ext[inside], exact[inside], rtol=0, atol=1e-12)
    assert np.allclose(r_ext, exact, rtol=0, atol=1e-9)
    with pytest.raises(ValueError, match="fall outside"):
        rg.regrid((x, y), go, f, method="multilinear", bounds="raise")
    with pytest.raises(ValueError, match="Unrecognized"):
        rg.regrid((x, y), go, f, method="multilinear", bounds="bogus")


@pytest.mark.gpu
def test_multilinear2d_fused_equals_weights_then_apply(env):
    """regrid(method='multilinear') (fused, no triplets) == weights() + regrid_from_weights() bit for bit,
    with batch axes; and the saved layout is the reference's (sorted, unique pairs, 4 per point)."""
    rg, torch, dev, oracle = env
    x, y = cases.curvilinear(19, 23, distort=0.01)
    go = cases.rectilinear_over(x, y, 15, 17, shrink=0.5)
    vals = np.random.default_rng(3).random((3, 2, 19, 23))
    fused = rg.regrid((x, y), go, vals, axis_input=(-2, -1), axis_output=(-2, -1), method="multilinear")
    W, shape_in, shape_out = rg.weights((x, y), go, method="multilinear")
    assert shape_in == (19, 23) and shape_out == (15, 17) and W.shape == ()
    ii, io, v = W[()]
    assert ii.dtype == np.int64 and io.dtype == np.int64 and v.dtype == np.float64 and ii.size == 4 * 15 * 17
    key = ii * (15 * 17) + io
    assert (np.diff(key) > 0).all()
    two = rg.regrid_from_weights(W, shape_in, shape_out, vals, axis_input=(-2, -1), axis_output=(-2, -1))
    assert fused.shape == (3, 2, 15, 17)
    assert np.array_equal(fused, two)
    ref = oracle.regrid_from_weights(ii, io, v, vals.reshape(6, -1), 15 * 17).reshape(3, 2, 15, 17)
    assert np.array_equal(fused, ref)


@pytest.mark.gpu
def test_multilinear2d_config5_properties(env):
    """Config-5 shape at reduced size (1024^2 vertices -> 2048^2 points): a field linear in (x, y) is
    reproduced to 1e-10 at every inside point, weights sum to 1."""
    rg, torch, dev, oracle = env
    n, m = 1024, 2048
    gi, _ = cases.benchmark_family(n, distorted=True)
    go = cases.rectilinear_over(gi[0], gi[1], m, m, shrink=0.7)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    x, y, px, py = t(gi[0]), t(gi[1]), t(go[0].reshape(-1)), t(go[1].reshape(-1))
    idx4, w4, n_out = rg.device.multilinear2d_weights(x, y, px, py, "nan")
    ok = ~torch.isnan(w4[:, 0])
    assert int(n_out.item()) == int((~ok).sum().item())
    assert float((w4[ok].sum(dim=1) - 1).abs().max()) < 1e-12
    f = (0.25 + 1.5 * x - 0.75 * y).reshape(1, -1)
    out = rg.device.ell4_apply(idx4, w4, f)[0]
    exact = 0.25 + 1.5 * px - 0.75 * py
    assert float((out[ok] - exact[ok]).abs().max()) < 1e-10
