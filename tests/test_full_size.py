"""
GPU parity at BASELINE.json's FULL sizes (``-m gpu``): the CUDA path against the CPU oracle on the same inputs,
bit for bit -- not only properties.  The oracle (OpenMP port of the reference, itself pinned to the reference's
goldens in ``test_oracle_golden.py``) builds a 2048^2-cell grid pair in ~10-30 s on the box's host cores.

Covers the size-dependent branches the small golden cases never reach: the piece cache overflowing
(``kPieceNone`` re-walks), buckets longer than the shared-memory sort capacity, > 64 fragments per input cell,
int32 cell ids near 2^22, tiles of the staged apply whose footprint does not fit.
"""

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def rg():
    import regridding_b200

    return regridding_b200


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _check_build_and_apply(rg, dev, oracle, gi, co, frames=4, seed=0):
    oracle.set_num_threads()
    dw = rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], device=dev)
    ii, io, v = dw.to_host()
    raw = oracle.weights_conservative_2d(gi, co)
    assert dw.stats["fragments"] == raw[2].size, "raw fragment count differs from the oracle's emission"
    oi, oo, ov = oracle.coalesce(*raw)
    del raw
    assert ii.size == oi.size
    assert np.array_equal(ii, oi) and np.array_equal(io, oo), "index structure differs from the oracle"
    assert np.array_equal(v, ov), f"weights differ: max rel {np.max(np.abs(v - ov) / np.abs(ov))}"
    assert dw.stats["repaired_segments"] >= 0 and dw.stats["unknown_guesses"] >= 0
    shape_in = (gi[0].shape[0] - 1, gi[0].shape[1] - 1)
    shape_out = (co[0].shape[0] - 1, co[0].shape[1] - 1)
    vals = np.random.default_rng(seed).random((frames, dw.n_in))
    ref = oracle.regrid_from_weights(oi, oo, ov, vals, dw.n_out)
    plan = dw.plan(shape_in, shape_out)
    out = rg.device.apply_planned(plan, T(vals, dev)).cpu().numpy()
    assert np.array_equal(out, ref), "staged apply differs from the oracle"
    out = rg.device.apply_csr(dw.csr(), T(vals, dev)).cpu().numpy()
    assert np.array_equal(out, ref), "generic apply differs from the oracle"
    return dw, plan


def test_config3_full_size_bit_exact_vs_oracle(rg, dev, oracle):
    """BASELINE config 3: distorted curvilinear 2049^2 vertices -> rectilinear 2049^2, seed-42 jitter."""
    n = 2049
    gi, go = cases.benchmark_family(n, distorted=True)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    dw, plan = _check_build_and_apply(rg, dev, oracle, gi, co, frames=4)
    assert dw.nnz > 14_000_000 and dw.stats["fragments"] > 59_000_000
    assert plan.n_generic_tiles <= plan.n_tiles // 1000


def test_config4_one_frame_full_size_bit_exact_vs_oracle(rg, dev, oracle):
    """BASELINE config 4: frame f carries its own grid (theta = 0.4 + 0.002 f, distortion phase f)."""
    n, f = 2049, 37
    gi, go = cases.benchmark_family(n, distorted=True, angle=0.4 + 0.002 * f, phase=float(f))
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    _check_build_and_apply(rg, dev, oracle, gi, co, frames=2, seed=f)


@pytest.mark.parametrize("n_in,n_out", [(385, 7), (7, 385), (1025, 33), (33, 1025), (300, 1100), (2049, 700)])
def test_extreme_resolution_ratios_bit_exact_vs_oracle(rg, dev, oracle, n_in, n_out):
    """Strong coarsening / refinement: thousands of fragments per input cell (long buckets, global-memory sort),
    sweep segments crossing dozens of static cells (piece-cache overflow), apply rows with thousands of entries
    and footprints far beyond a staged tile (generic tiles)."""
    gi, _ = cases.benchmark_family(n_in, distorted=True)
    _, go = cases.benchmark_family(n_in, n_out, n_out + 2, distorted=True)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    _check_build_and_apply(rg, dev, oracle, gi, co, frames=3)


def test_config5_locate_full_size_vs_oracle(rg, dev, oracle):
    """BASELINE config 5: 4096^2 curvilinear vertices, output points of the 8192^2 rectilinear grid over the
    FULL bounding box (so ~40 % fall outside); 120 000 sampled points against the reference's secant locator,
    200 of them against its exhaustive brute-force locator (lowest-index containing cell or the sentinel)."""
    oracle.set_num_threads()
    n, m = 4096, 8192
    gi, _ = cases.benchmark_family(n, distorted=True)
    xo = np.linspace(gi[0].min(), gi[0].max(), m)
    yo = np.linspace(gi[1].min(), gi[1].max(), m)
    rng = np.random.default_rng(5)
    P = 120_000
    a, b = rng.integers(0, m, P), rng.integers(0, m, P)
    # regular sub-lattice rows too (the walk locator seeds each point from its neighbour): whole output rows
    rows = np.array([0, 1, 4095, 4096, 8190, 8191])
    px = np.concatenate([xo[a], np.repeat(xo[rows], m)])
    py = np.concatenate([yo[b], np.tile(yo, rows.size)])
    xg, yg = T(gi[0], dev), T(gi[1], dev)
    got = rg.device.find_indices_2d(xg, yg, T(px, dev), T(py, dev), -1).cpu().numpy()
    want = oracle.index_of_points(gi[0], gi[1], px, py, -1, "secant")
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]
    outside = int((want < 0).sum())
    assert 0.2 * px.size < outside < 0.6 * px.size
    k = rng.choice(P, 200, replace=False)
    brute = oracle.index_of_points(gi[0], gi[1], px[k], py[k], -1, "brute")
    assert np.array_equal(got[k], brute)
    # the full 8192^2 lattice through the public API: every sampled point must agree with the flat result
    ri, rj = rg.find_indices(gi, np.meshgrid(xo, yo, indexing="ij"), fill_value=-1)
    flat = np.where(ri < 0, -1, ri * (n - 1) + rj)
    assert np.array_equal(flat[a, b], got[:P])
    for q, r in enumerate(rows):
        assert np.array_equal(flat[r], got[P + q * m:P + (q + 1) * m])
