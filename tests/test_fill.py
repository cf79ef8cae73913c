"""fill(method="gauss_seidel") -- SURVEY section 8 row f4.

Golden vectors come from the reference itself (tests/golden/make_golden_fill.py ran regridding.fill through
Numba in the build container).  CPU: the C oracle restatement against the goldens, bit for bit.  GPU: the CUDA
path through the C ABI against the goldens and against the oracle on larger seeded inputs, bit for bit (the
relaxation is fp64 with a fixed operation order; tolerance written below: exact equality)."""

import pathlib

import numpy as np
import pytest

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden" / "golden_fill.npz"


def _cases():
    g = np.load(GOLDEN)
    for name in sorted({k.split("/")[0] for k in g.files}):
        if name == "nan_array_guess":
            yield name, dict(a=g[name + "/a"], where=None, axis=(-2, -1), guess=g[name + "/guess_array"],
                             num_iterations=13), g[name + "/result"]
            continue
        ax = g[name + "/axis"]
        gu = float(g[name + "/guess"])
        yield name, dict(a=g[name + "/a"], where=g[name + "/where"], axis=None if ax[0] == -99 else tuple(int(x) for x in ax),
                         guess=None if np.isnan(gu) else gu, num_iterations=int(g[name + "/iters"])), g[name + "/result"]


def test_oracle_fill_matches_reference_goldens():
    from oracle import oracle

    n = 0
    for name, kw, expect in _cases():
        got = oracle.fill(**kw)
        assert np.array_equal(got, expect), name
        n += 1
    assert n == 9


def test_oracle_fill_strict_mode_is_the_noise_floor():
    """Plain IEEE evaluation of the source expression (NUMBA_DISABLE_JIT) differs from the JIT in the last bits."""
    from oracle import oracle

    name, kw, expect = next(c for c in _cases() if c[0] == "odd_odd")
    oracle.set_mode("strict")
    try:
        got = oracle.fill(**kw)
    finally:
        oracle.set_mode("jit")
    assert not np.array_equal(got, expect)
    assert np.allclose(got, expect, rtol=0, atol=1e-14)


@pytest.mark.gpu
def test_fill_matches_reference_goldens():
    import regridding_b200 as rg

    for name, kw, expect in _cases():
        got = rg.fill(method="gauss_seidel", **kw)
        assert got.shape == expect.shape and got.dtype == np.float64
        assert np.array_equal(got, expect), name


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 63, 65), (2, 128, 96), (1, 257, 31)])
def test_fill_matches_oracle_on_larger_inputs(shape):
    import regridding_b200 as rg
    from oracle import oracle

    rng = np.random.default_rng(sum(shape))
    a = rng.random(shape)
    where = np.zeros(shape, dtype=bool)
    for _ in range(12):  # blocks of missing cells, some across the periodic wrap
        t = rng.integers(0, shape[0])
        j, i = rng.integers(0, shape[1]), rng.integers(0, shape[2])
        jj = (j + np.arange(9)) % shape[1]
        ii = (i + np.arange(7)) % shape[2]
        where[t][np.ix_(jj, ii)] = True
    where |= rng.random(shape) < 0.05
    got = rg.fill(a, where=where, axis=(-2, -1), num_iterations=40)
    ref = oracle.fill(a, where=where, axis=(-2, -1), num_iterations=40)
    assert np.array_equal(got, ref)
    # the reference's own test properties (_fill_test.py:34-58)
    assert np.all(np.isfinite(got)) and np.array_equal(got[~where], a[~where]) and np.all(got[where] != 0)


@pytest.mark.gpu
def test_fill_api_behaviour():
    import regridding_b200 as rg

    a = np.random.default_rng(1).random((6, 7))
    a[2:4, 3:6] = np.nan
    b = rg.fill(a)  # NaN mask, all axes, median guess, 100 iterations
    assert np.isfinite(b).all() and np.array_equal(b[~np.isnan(a)], a[~np.isnan(a)])
    assert np.isnan(a).any()  # the input is not modified
    with pytest.raises(ValueError, match="interpolation axes"):
        rg.fill(np.zeros((3, 4, 5)), where=np.zeros((3, 4, 5), bool), axis=(0, 1, 2))
    with pytest.raises(ValueError, match="Unrecognized method"):
        rg.fill(a, method="bogus")
    same = rg.fill(a, where=np.zeros(a.shape, bool), num_iterations=3)  # nothing to fill
    assert np.array_equal(same, a, equal_nan=True)
