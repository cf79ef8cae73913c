"""
GPU parity tests (``-m gpu``): the CUDA path, called through the C ABI
(``regridding_b200._device`` = ctypes over ``libregrid_b200.so``) and through the
reference-compatible public API, against

  * the CPU oracle on the same seeded inputs (bit-exact: indices AND weights, because the
    kernels reproduce the reference's floating-point contraction pattern),
  * the committed golden vectors = outputs of the reference itself,
  * size-independent properties at BASELINE.json's full sizes.

Bars (north_star): triplet index sets bit-exact; weights within 1e-12 relative -- we hold
them to bit equality; resampled sums within 1e-10 relative -- also held to bit equality.
"""

import pickle

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


# ---------------------------------------------------------------------------
# 2D conservative build + apply
# ---------------------------------------------------------------------------


@pytest.mark.parametrize("name", cases.CASES_2D)
def test_build2d_bit_exact_vs_oracle_and_reference(rg, dev, oracle, golden, name):
    gi, go, w = cases.case_2d(name)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    dw = rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], w, device=dev)
    ii, io, v = dw.to_host()
    oi, oo, ov = oracle.coalesce(*oracle.weights_conservative_2d(gi, co, w))
    assert np.array_equal(ii, oi) and np.array_equal(io, oo)
    assert np.array_equal(v, ov), f"max rel {np.max(np.abs(v - ov) / np.abs(ov))}"
    assert dw.stats["fragments"] == int(golden[f"c2d/{name}/raw_n"])
    # the reference itself
    assert cases.sha(ii, io, v) == str(golden[f"c2d/{name}/final_sha"])
    # apply: three frames sharing the weights
    shape_in, shape_out = tuple(golden[f"c2d/{name}/shape_in"]), tuple(golden[f"c2d/{name}/shape_out"])
    vals = np.random.default_rng(0).random((3, *shape_in))
    out = rg.device.apply_csr(dw.csr(), T(vals.reshape(3, -1), dev)).cpu().numpy()
    ref = oracle.regrid_from_weights(oi, oo, ov, vals.reshape(3, -1), dw.n_out)
    assert np.array_equal(out, ref)
    assert cases.sha(out.reshape(3, *shape_out)) == str(golden[f"c2d/{name}/apply_sha"])
    # the shared-memory staged (planned) apply gives the same bits, incl. frame counts that do not fill a sub-block
    plan = dw.plan(shape_in, shape_out)
    for F in (3, 16, 37):
        vals_f = np.random.default_rng(F).random((F, dw.n_in))
        a = rg.device.apply_csr(dw.csr(), T(vals_f, dev))
        b = rg.device.apply_planned(plan, T(vals_f, dev))
        assert torch.equal(a, b), (name, F, plan.n_generic_tiles, plan.n_tiles)


def test_grid_area_bit_exact(rg, dev, oracle, golden):
    for name in ("fam40", "coarsen", "flipx"):
        g, _, _ = cases.case_2d(name)
        a = rg.device.grid_area(*g, device=dev).cpu().numpy()
        assert np.array_equal(a, golden[f"prim/volume_{name}"])
        assert np.array_equal(a, oracle.grid_volume(*g))


@pytest.mark.parametrize("n,mx,my", [(301, 280, 333), (513, 513, 513)])
def test_build2d_vs_oracle_benchmark_family(rg, dev, oracle, n, mx, my):
    """benchmarks/regrid.py family (libm sin/cos inputs, so oracle only -- no golden)."""
    gi, go = cases.benchmark_family(n, mx, my, distorted=True)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    dw = rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], device=dev)
    oi, oo, ov = oracle.coalesce(*oracle.weights_conservative_2d(gi, co))
    ii, io, v = dw.to_host()
    assert np.array_equal(ii, oi) and np.array_equal(io, oo) and np.array_equal(v, ov)


def test_build2d_band_partition_concatenates(rg, dev):
    """Input-cell bands (the multi-GPU partition) concatenate to the full build, bit for bit."""
    gi, go, _ = cases.case_2d("dist129")
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    full = rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], device=dev).to_host()
    ncx, ncy = gi[0].shape[0] - 1, gi[0].shape[1] - 1
    parts = []
    rows = [0, 17, 64, 65, ncx]
    for a, b in zip(rows[:-1], rows[1:]):
        parts.append(rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], cell_band=(a * ncy, b * ncy),
                                                device=dev).to_host())
    for k in range(3):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), full[k])


@pytest.mark.parametrize("name,world_size", [("dist129", 2), ("dist129", 3), ("dist129", 8), ("fam100", 5),
                                             ("coarsen", 4), ("refine", 4), ("inner", 3), ("flipx", 2),
                                             ("curv2curv", 3), ("winput", 2)])
def test_build2d_band_build_equals_full(rg, dev, name, world_size):
    """Exchange-free band build (rg_build2d_band; every rank walks only the segments that can reach its band of input
    rows, all ranks played on one GPU): the concatenated bands equal the single-GPU build bit for bit, and no chain
    of walk states needed the fallback."""
    from regridding_b200 import _parallel

    gi, go, w = cases.case_2d(name)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    full = rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], w, device=dev)
    ncx = gi[0].shape[0] - 1
    parts, nfrag = [], 0
    for r in range(world_size):
        lo, hi = _parallel.shard_range(ncx, r, world_size)
        bb = rg.device.build2d_band_enqueue(T(gi[0], dev), T(gi[1], dev), T(co[0], dev), T(co[1], dev),
                                            None if w is None else T(w, dev), lo, hi, device=dev)
        dw, status = bb.finish()
        if status == "capacity":
            dw, status = rg.device.build2d_band_enqueue(T(gi[0], dev), T(gi[1], dev), T(co[0], dev), T(co[1], dev),
                                                        None if w is None else T(w, dev), lo, hi, device=dev).finish()
        assert status == "ok", (name, world_size, r, status)
        parts.append(dw)
        nfrag += dw.stats["fragments"]
        # the second build of a shape knows the longest bucket and walks ONCE (rg_build2d_band_onewalk): same band
        again = rg.device.build2d_band_enqueue(T(gi[0], dev), T(gi[1], dev), T(co[0], dev), T(co[1], dev),
                                               None if w is None else T(w, dev), lo, hi, device=dev)
        assert again.bucket_capacity > 0, (name, r)
        dw2, status2 = again.finish()
        assert status2 == "ok", (name, world_size, r, status2)
        assert torch.equal(dw2.indices_input, dw.indices_input) and torch.equal(dw2.indices_output, dw.indices_output)
        assert torch.equal(dw2.values, dw.values)
        assert dw2.stats["fragments"] == dw.stats["fragments"]
    assert nfrag == full.stats["fragments"]
    assert torch.equal(torch.cat([p.indices_input for p in parts]), full.indices_input)
    assert torch.equal(torch.cat([p.indices_output for p in parts]), full.indices_output)
    assert torch.equal(torch.cat([p.values for p in parts]), full.values)
    # the convenience wrapper (with its fallback path) gives the same band
    lo, hi = _parallel.shard_range(ncx, world_size - 1, world_size)
    one = rg.device.build_weights_2d_band(gi[0], gi[1], co[0], co[1], w, row_band=(lo, hi), device=dev)
    assert torch.equal(one.values, parts[-1].values) and torch.equal(one.indices_output, parts[-1].indices_output)


def test_build2d_band_onewalk_bucket_overflow_falls_back(rg, dev):
    """One-walk band build whose fixed-capacity buckets are too small (capacity learned on other coordinates): it must
    report "capacity", the repeat must take two walks and be right, and the shape must stay with two walks."""
    from regridding_b200 import _parallel, _device

    gi, go, _ = cases.case_2d("dist129")
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    ncx, ncy = gi[0].shape[0] - 1, gi[0].shape[1] - 1
    lo, hi = _parallel.shard_range(ncx, 1, 3)
    t = [T(a, dev) for a in (gi[0], gi[1], co[0], co[1])]
    full = rg.device.build_weights_2d(*t, device=dev)
    sel = (full.indices_input >= lo * ncy) & (full.indices_input < hi * ncy)
    dw, status = rg.device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
    if status == "capacity":
        dw, status = rg.device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
    assert status == "ok"
    key = (*gi[0].shape, *co[0].shape, lo, hi)
    _device._band_onewalk_failed.discard(key)
    fcap, ncap, bcap = _device._band_caps[key]
    assert bcap > 0
    _device._band_caps[key] = (fcap, ncap, 2)   # far too small
    bb = rg.device.build2d_band_enqueue(*t, None, lo, hi, device=dev)
    assert bb.bucket_capacity == 2
    none, status = bb.finish()
    assert none is None and status == "capacity"
    bb = rg.device.build2d_band_enqueue(*t, None, lo, hi, device=dev)
    assert bb.bucket_capacity == 0
    dw2, status = bb.finish()
    assert status == "ok"
    assert torch.equal(dw2.indices_output, full.indices_output[sel]) and torch.equal(dw2.values, full.values[sel])
    assert rg.device.build2d_band_enqueue(*t, None, lo, hi, device=dev).bucket_capacity == 0   # stays with two walks
    _device._band_onewalk_failed.discard(key)


def test_build2d_band_unverifiable_states_fall_back(rg, dev, monkeypatch):
    """A band whose walk states do not all verify must say so ("mismatch") and the wrapper must then rebuild THAT band
    with the sequentially verified banded build, on its own (no other rank is involved).  Forced through the test hook
    RG_BAND_FORCE_MISMATCH: grids with sweep vertices within rounding error of a static cell edge -- what the exact check
    of a run's first state rejects -- are outside the domain of the sweep itself (no walk terminates on them, here as in
    the reference, which perturbs its output grid for that reason)."""
    from regridding_b200 import _parallel

    gi, go, _ = cases.case_2d("dist129")
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    ncx, ncy = gi[0].shape[0] - 1, gi[0].shape[1] - 1

    def check(xo, yo, world_size, expect=None):
        full = rg.device.build_weights_2d(gi[0], gi[1], xo, yo, device=dev)
        seen = []
        for r in range(world_size):
            lo, hi = _parallel.shard_range(ncx, r, world_size)
            t = [T(a, dev) for a in (gi[0], gi[1], xo, yo)]
            dw, status = rg.device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
            if status == "capacity":
                dw, status = rg.device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
            seen.append(status)
            sel = (full.indices_input >= lo * ncy) & (full.indices_input < hi * ncy)
            if status == "ok":
                assert torch.equal(dw.indices_output, full.indices_output[sel]) and torch.equal(dw.values, full.values[sel]), r
            one = rg.device.build_weights_2d_band(gi[0], gi[1], xo, yo, None, row_band=(lo, hi), device=dev)
            assert torch.equal(one.indices_input, full.indices_input[sel]), r
            assert torch.equal(one.indices_output, full.indices_output[sel]), r
            assert torch.equal(one.values, full.values[sel]), r
        if expect is not None:
            assert all(s == expect for s in seen), seen
        return seen

    monkeypatch.setenv("RG_BAND_FORCE_MISMATCH", "1")
    check(co[0], co[1], 3, expect="mismatch")
    monkeypatch.delenv("RG_BAND_FORCE_MISMATCH")
    check(co[0], co[1], 3, expect="ok")


def test_build2d_band_rebuilt_with_grids_updated_in_place(rg, dev):
    """The band build re-run on the SAME buffers (scratch, learned capacities) while the output grid is rewritten in
    place between the calls (the per-frame grids of _weights_conservative.py:110-139): every rebuild must equal a fresh
    single-GPU build of the current coordinates -- nothing of a previous build may leak through the reused workspace."""
    from regridding_b200 import _parallel

    n = 161
    gi, _ = cases.benchmark_family(n, distorted=True)
    xi, yi = T(gi[0], dev), T(gi[1], dev)
    xo = torch.empty((n, n), dtype=torch.float64, device=dev)
    yo = torch.empty((n, n), dtype=torch.float64, device=dev)
    lo, hi = _parallel.shard_range(n - 1, 1, 3)
    ncy = n - 1
    kept = None   # the previous result stays alive for one round, as in `dw = build(...)` loops: two buffer sets alternate
    for f in range(8):
        _, go = cases.benchmark_family(n, distorted=True, angle=0.4 + 0.01 * (f // 2), phase=float(f // 2))
        co = cases.perturb_like_reference(go, (-1, -2), 42)
        xo.copy_(torch.from_numpy(co[0]))
        yo.copy_(torch.from_numpy(co[1]))
        dw, status = rg.device.build2d_band_enqueue(xi, yi, xo, yo, None, lo, hi, device=dev).finish()
        if status == "capacity":
            dw, status = rg.device.build2d_band_enqueue(xi, yi, xo, yo, None, lo, hi, device=dev).finish()
        assert status == "ok", (f, status)
        full = rg.device.build_weights_2d(xi, yi, xo, yo, device=dev)
        sel = (full.indices_input >= lo * ncy) & (full.indices_input < hi * ncy)
        assert torch.equal(dw.indices_input, full.indices_input[sel]), f
        assert torch.equal(dw.indices_output, full.indices_output[sel]), f
        assert torch.equal(dw.values, full.values[sel]), f
        kept = dw
    assert kept is not None


@pytest.mark.parametrize("name,world_size", [("dist129", 2), ("dist129", 3), ("dist129", 8), ("fam100", 5)])
def test_build2d_line_sharded_equals_full(rg, dev, name, world_size):
    """Line-sharded build (every rank walks 1/W of the sweep lines, fragments exchanged to the band owners,
    all ranks played on one GPU): the concatenated bands equal the single-GPU build bit for bit."""
    from regridding_b200 import _parallel

    gi, go, _ = cases.case_2d(name)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    full = rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], device=dev)
    bands = _parallel.build_weights_2d_sharded_local(gi[0], gi[1], co[0], co[1], world_size=world_size, device=dev)
    assert len(bands) == world_size
    assert sum(b.stats["fragments"] for b in bands) == full.stats["fragments"]
    assert sum(b.stats["fragments_walked"] for b in bands) == full.stats["fragments"]
    got = [torch.cat([getattr(b, k) for b in bands]) for k in ("indices_input", "indices_output", "values")]
    assert torch.equal(got[0], full.indices_input) and torch.equal(got[1], full.indices_output)
    assert torch.equal(got[2], full.values)
    bounds = _parallel.band_bounds(gi[0].shape[0] - 1, gi[0].shape[1] - 1, world_size)
    for r, b in enumerate(bands):
        if b.nnz:
            assert bounds[r] <= int(b.indices_input.min()) and int(b.indices_input.max()) < bounds[r + 1]


def test_build2d_line_sharded_weights_input_and_more_ranks_than_blocks(rg, dev):
    """weights_input rides along; with more ranks than 32-line blocks some ranks walk nothing."""
    from regridding_b200 import _parallel

    gi, go, _ = cases.case_2d("fam100")
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    w = np.random.default_rng(5).random((gi[0].shape[0] - 1, gi[0].shape[1] - 1)) + 0.5
    full = rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], w, device=dev)
    bands = _parallel.build_weights_2d_sharded_local(gi[0], gi[1], co[0], co[1], w, world_size=7, device=dev)
    for k in ("indices_input", "indices_output", "values"):
        assert torch.equal(torch.cat([getattr(b, k) for b in bands]), getattr(full, k))
    # coarse output grid: 2 blocks of lines per pass for 7 ranks
    gi2 = cases.curvilinear(40, 33)
    go2 = cases.rectilinear_over(gi2[0], gi2[1], 21, 19)
    full = rg.device.build_weights_2d(gi2[0], gi2[1], go2[0], go2[1], device=dev)
    bands = _parallel.build_weights_2d_sharded_local(gi2[0], gi2[1], go2[0], go2[1], world_size=7, device=dev)
    for k in ("indices_input", "indices_output", "values"):
        assert torch.equal(torch.cat([getattr(b, k) for b in bands]), getattr(full, k))


def test_build2d_disjoint_and_tiny_grids(rg, dev, oracle):
    # no overlap at all -> empty weights
    gi = cases.curvilinear(9, 7)
    go = cases.rectilinear_over(gi[0] + 10.0, gi[1], 5, 6)
    dw = rg.device.build_weights_2d(gi[0], gi[1], go[0], go[1], device=dev)
    assert dw.nnz == 0
    out = rg.device.apply_csr(dw.csr(), torch.ones((2, dw.n_in), dtype=torch.float64, device=dev))
    assert out.shape == (2, dw.n_out) and float(out.abs().sum()) == 0.0
    # single cell onto single cell (first case of _weights_conservative_2d_test.py:27-44)
    bx, by = np.meshgrid(np.linspace(-1, 1, 2), np.linspace(-1, 1, 2), indexing="ij")
    dw = rg.device.build_weights_2d(bx, by, 2 * bx + 1e-6, 2 * by + 1e-6, device=dev)
    oi, oo, ov = oracle.coalesce(*oracle.weights_conservative_2d((bx, by), (2 * bx + 1e-6, 2 * by + 1e-6)))
    ii, io, v = dw.to_host()
    assert np.array_equal(ii, oi) and np.array_equal(io, oo) and np.array_equal(v, ov)
    assert np.allclose(v.sum(), 1.0, rtol=1e-3)


def test_config3_full_size_properties(rg, dev):
    """2048x2048 cells -> 2048x2048 cells (BASELINE.json config 3): structure and conservation."""
    n = 2049
    gi, go = cases.benchmark_family(n, distorted=True)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    dw = rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], device=dev)
    ii, io, v = dw.indices_input, dw.indices_output, dw.values
    assert dw.stats["repaired_segments"] >= 0
    # sorted by (input, output), pairs unique
    key = ii * dw.n_out + io
    assert bool((key[1:] > key[:-1]).all())
    assert int(ii.min()) >= 0 and int(ii.max()) < dw.n_in and int(io.min()) >= 0 and int(io.max()) < dw.n_out
    # the output grid covers the input grid, so every input cell is fully redistributed:
    # the weights of each input cell sum to 1 (up to eps * (r/h)^2)
    colsum = torch.zeros(dw.n_in, dtype=torch.float64, device=dev).index_add_(0, ii, v)
    assert float((colsum - 1).abs().max()) < 1e-8
    # apply: constant field in -> area-weighted constant out wherever the output cell is covered;
    # linearity: A(a x + b y) = a A(x) + b A(y) up to rounding
    F = 4
    x = torch.rand((F, dw.n_in), dtype=torch.float64, device=dev)
    y = torch.rand((F, dw.n_in), dtype=torch.float64, device=dev)
    csr = dw.csr()
    ax, ay = rg.device.apply_csr(csr, x), rg.device.apply_csr(csr, y)
    axy = rg.device.apply_csr(csr, (2.0 * x + 0.5 * y).contiguous())
    assert float((axy - (2.0 * ax + 0.5 * ay)).abs().max()) < 1e-9
    plan = dw.plan((n - 1, n - 1), (n - 1, n - 1))
    assert plan.n_generic_tiles <= plan.n_tiles // 1000  # a few boundary tiles may take the generic kernel
    assert torch.equal(rg.device.apply_planned(plan, x), ax)
    # CSR apply == straightforward COO scatter-add of the public triplets (different summation order)
    ref = torch.zeros((F, dw.n_out), dtype=torch.float64, device=dev)
    ref.index_add_(1, io, x[:, ii] * v)
    assert float((ax - ref).abs().max() / ref.abs().max()) < 1e-12
    # total mass: sum_out = sum_in of (value * column sum)
    assert abs(float(ax.sum()) - float((x * colsum).sum())) / float(ax.sum()) < 1e-10


# ---------------------------------------------------------------------------
# public API (reference signatures)
# ---------------------------------------------------------------------------


def test_api_weights_layout_and_roundtrip(rg, golden):
    """regridding/_regrid/_tests/test_regrid_from_weights.py:38-117."""
    gi, go, _ = cases.case_2d("fam40")
    W = rg.weights(gi, go, method="conservative")
    weights, shape_in, shape_out = W
    assert weights.dtype == object and weights.shape == ()
    ii, io, v = weights[()]
    assert ii.ndim == 1 and ii.shape == io.shape == v.shape
    assert ii.dtype == np.int64 and io.dtype == np.int64 and v.dtype == np.float64
    assert shape_in == tuple(golden["c2d/fam40/shape_in"]) and shape_out == tuple(golden["c2d/fam40/shape_out"])
    assert np.array_equal(ii, golden["c2d/fam40/ii"]) and np.array_equal(io, golden["c2d/fam40/io"])
    assert np.array_equal(v, golden["c2d/fam40/v"])
    key = ii * (io.max() + 1) + io
    assert np.unique(key).size == key.size
    vals = np.random.default_rng(0).random((3, *shape_in))
    res = rg.regrid_from_weights(*W, vals)
    assert res.dtype == np.float64 and np.array_equal(res, golden["c2d/fam40/apply"])
    # pickle round trip gives bitwise-equal regrids
    res2 = rg.regrid_from_weights(*pickle.loads(pickle.dumps(W)), vals)
    assert np.array_equal(res, res2)
    # regrid == weights + regrid_from_weights, bitwise
    res3 = rg.regrid(gi, go, vals, method="conservative")
    assert np.array_equal(res, res3)
    # values_output supplied: filled in place
    buf = np.full_like(res, 7.0)
    res4 = rg.regrid_from_weights(*W, vals, values_output=buf)
    assert np.array_equal(buf, res) and np.shares_memory(res4, buf)
    with pytest.raises(ValueError):
        rg.regrid_from_weights(*W, vals, values_output=np.zeros((2, 2)))


def test_api_weights_input(rg, golden):
    gi, go, w = cases.case_2d("winput")
    W = rg.weights(gi, go, weights_input=w, method="conservative")
    assert cases.sha(*W[0][()]) == str(golden["c2d/winput/final_sha"])
    vals = np.random.default_rng(0).random((3, *W[1]))
    assert np.array_equal(rg.regrid_from_weights(*W, vals), golden["c2d/winput/apply"])


def test_api_batched_frames_and_seeds(rg, golden):
    """Per-slice grids (orthogonal axis 0) + the perturbation stream order and seeding rules
    (regridding/_weights/_weights_test.py:39-103)."""
    gi, go = cases.case_2d_batched()
    kw = dict(axis_input=(1, 2), axis_output=(1, 2), method="conservative")
    W, shape_in, shape_out = rg.weights(gi, go, **kw)
    assert W.shape == (3,) and shape_in == tuple(golden["c2d_batched/shape_in"])
    for f in range(3):
        assert np.array_equal(W[f][0], golden[f"c2d_batched/{f}/ii"])
        assert np.array_equal(W[f][1], golden[f"c2d_batched/{f}/io"])
        assert np.array_equal(W[f][2], golden[f"c2d_batched/{f}/v"])
    vals = np.random.default_rng(0).random(shape_in)
    res = rg.regrid_from_weights(W, shape_in, shape_out, vals, axis_input=(1, 2), axis_output=(1, 2))
    assert np.array_equal(res, golden["c2d_batched/apply"])
    assert np.array_equal(rg.regrid(gi, go, vals, **kw), golden["c2d_batched/apply"])
    # seeds
    W7, *_ = rg.weights(gi, go, seed=7, **kw)
    assert np.array_equal(W7[0][2], golden["c2d_batched/seed7_v0"])
    Wg, *_ = rg.weights(gi, go, seed=np.random.default_rng(7), **kw)
    assert np.array_equal(Wg[0][2], W7[0][2])
    Wn, *_ = rg.weights(gi, go, perturb=False, seed=123, **kw)
    assert np.array_equal(Wn[0][2], golden["c2d_batched/noperturb_v0"])
    assert np.array_equal(Wn[0][0], golden["c2d_batched/noperturb_ii0"])
    Wr1, *_ = rg.weights(gi, go, seed=None, **kw)
    Wr2, *_ = rg.weights(gi, go, seed=None, **kw)
    assert not (Wr1[0][2].shape == Wr2[0][2].shape and np.array_equal(Wr1[0][2], Wr2[0][2]))


def test_api_host_pipeline_chunks_and_trailing_orthogonal_axes(rg, golden, monkeypatch):
    """The host path streams chunks of orthogonal slices through a ring of device buffers
    (H2D / apply / D2H overlapped); any chunking gives the same bits.  Also the reference's
    in-place rule (rfw.py:120-154): with orthogonal axes TRAILING the caller's buffer is left
    zeroed and the result comes back in a fresh array."""
    from regridding_b200 import _regrid

    gi, go, _ = cases.case_2d("fam40")
    W = rg.weights(gi, go, method="conservative")
    F = 11
    vals = np.random.default_rng(0).random((F, *W[1]))
    whole = rg.regrid_from_weights(*W, vals)
    assert np.array_equal(whole[:3], golden["c2d/fam40/apply"])
    n_in = int(np.prod(W[1]))
    monkeypatch.setattr(_regrid, "_CHUNK_BYTES", 8 * n_in * 2)  # 2 frames per chunk -> 6 chunks, ring of 3
    buf = np.full((F, *W[2]), 5.0)
    chunked = rg.regrid_from_weights(*W, vals, values_output=buf)
    assert np.array_equal(chunked, whole) and np.shares_memory(chunked, buf)
    # per-slice weights (batched build): chunks never straddle two weights elements
    gib, gob = cases.case_2d_batched()
    kw = dict(axis_input=(1, 2), axis_output=(1, 2))
    Wb, sib, sob = rg.weights(gib, gob, method="conservative", **kw)
    vb = np.random.default_rng(0).random(sib)
    monkeypatch.setattr(_regrid, "_CHUNK_BYTES", 1)
    assert np.array_equal(rg.regrid_from_weights(Wb, sib, sob, vb, **kw), golden["c2d_batched/apply"])
    # 1D weights, orthogonal axis trailing: result in a fresh array, caller's buffer zeroed
    xin, xout, _ = cases.cases_1d()["spectra"]
    W1 = rg.weights((xin.T.copy(),), (xout.T.copy(),), axis_input=0, axis_output=0, method="conservative")
    v1 = np.random.default_rng(0).random(W1[1])
    lead = rg.regrid_from_weights(*rg.weights((xin,), (xout,), axis_input=-1, axis_output=-1, method="conservative"),
                                  v1.T.copy(), axis_input=-1, axis_output=-1)
    out1 = np.full(W1[2], 3.0)
    res1 = rg.regrid_from_weights(*W1, v1, values_output=out1, axis_input=0, axis_output=0)
    assert np.array_equal(res1, lead.T)
    assert not np.shares_memory(res1, out1) and not out1.any()


def test_api_shared_weights_over_frames_and_device_tensors(rg, dev, golden):
    """2D weights applied to (F, H, W) values: F is an orthogonal axis sharing the weights (rfw.py:108)."""
    gi, go, _ = cases.case_2d("fam40")
    W = rg.weights(gi, go, method="conservative")
    vals = np.random.default_rng(0).random((3, *W[1]))
    res = rg.regrid_from_weights(*W, vals)
    assert res.shape == (3, *W[2])
    one = np.stack([rg.regrid_from_weights(*W, vals[f]) for f in range(3)])
    assert np.array_equal(res, one)
    # trailing batch axis with explicit axes is a broadcast error, like the reference (rfw.py:109)
    with pytest.raises(ValueError):
        rg.regrid_from_weights(*W, np.moveaxis(vals, 0, -1), axis_input=(0, 1), axis_output=(0, 1))
    # scalar values
    s = rg.regrid_from_weights(*W, np.float64(2.0))
    assert s.shape == W[2]
    # device tensors in -> device tensor out, same bits
    rd = rg.regrid_from_weights(*W, torch.from_numpy(vals).to(dev))
    assert isinstance(rd, torch.Tensor) and rd.is_cuda and np.array_equal(rd.cpu().numpy(), res)


ANALYTIC_2D = "regridding/_weights/_weights_conservative_2d/_weights_conservative_2d_test.py:16-301"


def _analytic_cases():
    box_x = np.linspace(-1, 1, num=2)[..., np.newaxis]
    box_y = np.linspace(-1, 1, num=2)
    x = np.linspace(-1, 1, num=6)[..., np.newaxis]
    y = np.linspace(-1, 1, num=6)
    x2 = np.linspace(-1, 1, num=11)[..., np.newaxis]
    y2 = np.linspace(-1, 1, num=11)
    c90, s90 = 0.0, 1.0
    ones5 = np.ones((5, 5))
    A = np.random.RandomState(42).uniform(0, 10, size=(5, 5))
    out = [
        ((box_x, box_y), (2 * box_x, 2 * box_y), np.array([[1]]), None, np.array([[1]])),
        ((-box_x, -box_y), (2 * box_x, 2 * box_y), np.array([[1]]), None, np.array([[1]])),
        ((2 * box_x, 2 * box_y), (box_x, box_y), np.array([[1]]), None, np.array([[0.25]])),
        ((x, y), (x, y), A, None, A),
        ((x, y), (x2, y2), ones5, None, 0.25 * np.ones((10, 10))),
        ((x2, y2), (x, y), np.ones((10, 10)), None, 4 * ones5),
        ((x, y), (x * c90 - y * s90, x * s90 + y * c90), A, None,
         np.rot90(A, -1)),
        ((x, y), (x, y), ones5, 2 * ones5, 2 * ones5),
        ((x, y), (-x, -y), A, None, A[::-1, ::-1]),
        ((x * c90 - y * s90, x * s90 + y * c90), (x, y), A, None, np.rot90(A)),
    ]
    return out


@pytest.mark.parametrize("k", range(10))
def test_api_analytic_cases_like_reference_tests(rg, oracle, k):
    """Analytic expectations in the style of the reference's own 2D test module
    (output shifted by 1e-6, perturb=False, rtol=1e-3); cross-checked against the oracle."""
    ci, co, vals, w, expected = _analytic_cases()[k]
    co = (co[0] + 1e-6, co[1] + 1e-6)
    W = rg.weights(ci, co, weights_input=w, method="conservative", perturb=False)
    res = rg.regrid_from_weights(*W, vals)
    assert np.allclose(res, expected, rtol=1e-3)
    gi = tuple(np.ascontiguousarray(np.broadcast_to(c, np.broadcast(*ci).shape), dtype=float) for c in ci)
    go = tuple(np.ascontiguousarray(np.broadcast_to(c, np.broadcast(*co).shape), dtype=float) for c in co)
    wi = None if w is None else np.ascontiguousarray(w, dtype=float)
    oi, oo, ov = oracle.coalesce(*oracle.weights_conservative_2d(gi, go, wi))
    ii, io, v = W[0][()]
    assert np.array_equal(ii, oi) and np.array_equal(io, oo) and np.array_equal(v, ov)


# ---------------------------------------------------------------------------
# 1D conservative
# ---------------------------------------------------------------------------


@pytest.mark.parametrize("name", list(cases.cases_1d()))
def test_conservative_1d_vs_reference(rg, dev, oracle, golden, name):
    xin, xout, w = cases.cases_1d()[name]
    W = rg.weights((xin,), (xout,), axis_input=-1, axis_output=-1, weights_input=w, method="conservative")
    flat = W[0].reshape(-1)
    assert [e[2].size for e in flat] == list(golden[f"c1d/{name}/counts"])
    assert np.array_equal(np.concatenate([e[0] for e in flat]), golden[f"c1d/{name}/ii"])
    assert np.array_equal(np.concatenate([e[1] for e in flat]), golden[f"c1d/{name}/io"])
    assert np.array_equal(np.concatenate([e[2] for e in flat]), golden[f"c1d/{name}/v"])
    vals = np.random.default_rng(0).random(W[1])
    res = rg.regrid_from_weights(*W, vals, axis_input=-1, axis_output=-1)
    assert np.array_equal(res, golden[f"c1d/{name}/apply"])
    # raw emission order of the kernel call site == oracle walk
    ii, io, v, counts = rg.device.cons1d_batched(T(xin, dev), T(xout, dev), None if w is None else T(w, dev))
    raw = oracle.weights_conservative_1d_batched(xin, xout, w)
    for s, (ri, ro, rv) in enumerate(raw):
        c = int(counts[s])
        assert c == rv.size
        assert np.array_equal(ii[s, :c].cpu().numpy(), ri) and np.array_equal(io[s, :c].cpu().numpy(), ro)
        assert np.array_equal(v[s, :c].cpu().numpy(), rv)
    # fused regrid (weights never materialised) gives the same bits
    fused = rg.device.regrid1d_conservative(T(xin, dev), T(xout, dev), T(vals, dev), None if w is None else T(w, dev))
    assert np.array_equal(fused.cpu().numpy(), golden[f"c1d/{name}/apply"])
    # and so does regrid()
    if w is None:
        assert np.array_equal(rg.regrid((xin,), (xout,), vals, axis_input=-1, axis_output=-1, method="conservative"),
                              golden[f"c1d/{name}/apply"])


def test_conservative_1d_config2_shape_properties(rg, dev, oracle):
    """BASELINE.json config 2 at reduced S: 4096 bins, per-spectrum grids; fused == oracle; flux conserved."""
    rng = np.random.default_rng(0)
    S, n = 64, 4097
    base = np.linspace(4000.0, 7000.0, n)
    xin = base * (1 + 1e-4 * rng.standard_normal((S, 1))) + 0.3 * np.sin(base / 500 + rng.random((S, 1)))
    xout = np.linspace(4001.0, 6999.0, n) + 0.05 * rng.random((S, 1))
    vals = rng.random((S, n - 1))
    fused = rg.device.regrid1d_conservative(T(xin, dev), T(xout, dev), T(vals, dev)).cpu().numpy()
    for s in (0, 17, 63):
        tri = oracle.coalesce(*oracle.weights_conservative_1d(xin[s], xout[s]))
        ref = oracle.regrid_from_weights(*tri, vals[s:s + 1], n - 1)[0]
        assert np.array_equal(fused[s], ref)
        assert tri[2].size == 8189 or abs(tri[2].size - 8189) < 8


def test_conservative_1d_fused_paths_agree(rg, dev, oracle):
    """The fused 1D regrid has a shared-memory staged kernel (one spectrum on chip, bracketed searches) and a
    generic one for spectra too long for that; both must give the oracle's bits, also for non-uniform grids
    (bracket fall-back), descending grids and weights."""
    rng = np.random.default_rng(5)
    for n, m in ((700, 333), (9001, 8800)):  # staged / generic (n + m + n-1 doubles > 200 KB)
        S = 3
        xin = np.sort(rng.uniform(0.0, 100.0, (S, n)), axis=1)
        xin[1] = xin[1, ::-1]                       # one descending input grid
        xin[2] = np.linspace(0, 100, n) ** 1.0      # one uniform
        xout = np.sort(rng.uniform(-5.0, 105.0, (S, m)), axis=1)
        xout[2] = xout[2, ::-1]                     # one descending output grid
        vals = rng.random((S, n - 1))
        w = rng.random((S, n - 1))
        for ww in (None, w):
            fused = rg.device.regrid1d_conservative(T(xin, dev), T(xout, dev), T(vals, dev),
                                                    None if ww is None else T(ww, dev)).cpu().numpy()
            for s_ in range(S):
                tri = oracle.coalesce(*oracle.weights_conservative_1d(xin[s_], xout[s_], None if ww is None else ww[s_]))
                ii_, io_ = tri[0] % (n - 1), tri[1] % (m - 1)
                ref = oracle.regrid_from_weights(ii_, io_, tri[2], vals[s_:s_ + 1], m - 1)[0]
                assert np.array_equal(fused[s_], ref), (n, m, s_, ww is None)


def test_conservative_1d_streamed_kernel_equals_staged_and_oracle(rg, dev, oracle, monkeypatch):
    """The streamed kernel (persistent CTAs, two stages filled by bulk copies) against the staged one and the oracle:
    more spectra than CTAs (several iterations per CTA, both stages), odd and even row lengths (rows of odd length
    start 8 bytes off every other spectrum: skewed copies), mis-aligned base pointers (first / last spectrum staged by
    plain loads), descending and far-from-uniform grids (search fall-back), output cells beyond the input range."""
    rng = np.random.default_rng(11)
    for n, m, S in ((4097, 4097, 301), (1000, 1201, 310), (513, 64, 5), (2, 2, 3), (3, 700, 2)):
        xin = np.sort(rng.uniform(0.0, 100.0, (S, n)), axis=1)
        xin[::7] = np.linspace(0.0, 100.0, n) + 0.3 * np.sin(np.linspace(0, 20, n))      # smooth grids: guesses verify
        xin[1] = xin[1, ::-1].copy()                                                     # descending input grid
        xin[S - 1] = np.cumsum(rng.exponential(1.0, n)) ** 2                             # strongly non-uniform
        xout = np.sort(rng.uniform(-5.0, 105.0, (S, m)), axis=1)
        xout[2 % S] = xout[2 % S, ::-1].copy()                                           # descending output grid
        vals = rng.random((S, n - 1))
        import torch

        for shift in (0, 1):   # 1: every buffer starts 8 bytes off a 16-byte boundary
            def dev_of(a):
                flat = torch.empty(a.size + 2, dtype=torch.float64, device=dev)
                view = flat[shift:shift + a.size].view(a.shape)
                view.copy_(torch.from_numpy(a))
                return view
            a, b, c = dev_of(xin), dev_of(xout), dev_of(vals)
            monkeypatch.delenv("RG_NO_STREAM1D", raising=False)
            streamed = rg.device.regrid1d_conservative(a, b, c).cpu().numpy()
            monkeypatch.setenv("RG_NO_STREAM1D", "1")
            staged = rg.device.regrid1d_conservative(a, b, c).cpu().numpy()
            monkeypatch.delenv("RG_NO_STREAM1D", raising=False)
            assert np.array_equal(streamed, staged), (n, m, S, shift)
        for s_ in sorted({0, 1, 2 % S, 7 % S, S // 2, S - 1}):
            tri = oracle.coalesce(*oracle.weights_conservative_1d(xin[s_], xout[s_]))
            ref = oracle.regrid_from_weights(tri[0] % (n - 1), tri[1] % (m - 1), tri[2], vals[s_:s_ + 1], m - 1)[0]
            assert np.array_equal(streamed[s_], ref), (n, m, s_)


def test_regrid_1d_fused_fast_path_equals_weights_path(rg):
    """regrid(method="conservative") along one axis takes the fused kernel (no weights materialised);
    it must return what weights() + regrid_from_weights() return, for any axis position, broadcast
    coordinates and a caller-supplied output buffer."""
    rng = np.random.default_rng(3)
    n, m = 41, 29
    xin = np.sort(rng.uniform(0, 10, (3, n, 2)), axis=1)
    xout = np.sort(rng.uniform(-1, 11, (1, m, 2)), axis=1)          # broadcast along the first orthogonal axis
    vals = rng.random((3, n - 1, 2))
    kw = dict(axis_input=1, axis_output=1, method="conservative")
    W = rg.weights((xin,), (xout,), **kw)
    ref = rg.regrid_from_weights(*W, vals, axis_input=1, axis_output=1)
    got = rg.regrid((xin,), (xout,), vals, **kw)
    assert got.shape == (3, m - 1, 2) and np.array_equal(got, ref)
    # resampled axis last + output buffer: filled in place
    xin2, xout2, vals2 = (np.ascontiguousarray(np.moveaxis(a, 1, -1)) for a in (xin, np.broadcast_to(xout, (3, m, 2)), vals))
    buf = np.full((3, 2, m - 1), 9.0)
    got2 = rg.regrid((xin2,), (xout2,), vals2, values_output=buf, axis_input=-1, axis_output=-1, method="conservative")
    assert np.shares_memory(got2, buf) and np.array_equal(np.moveaxis(buf, -1, 1), ref)
    # resampled axis first + output buffer: the reference leaves the buffer zeroed and returns a fresh array
    buf3 = np.full((m - 1, 3, 2), 9.0)
    got3 = rg.regrid((np.moveaxis(xin, 1, 0),), (np.moveaxis(np.broadcast_to(xout, (3, m, 2)), 1, 0),),
                     np.moveaxis(vals, 1, 0), values_output=buf3, axis_input=0, axis_output=0, method="conservative")
    assert not buf3.any() and np.array_equal(np.moveaxis(got3, 0, 1), ref)
    with pytest.raises(ValueError):
        rg.regrid((xin,), (xout,), vals, values_output=np.zeros((2, 2)), **kw)


# ---------------------------------------------------------------------------
# find_indices
# ---------------------------------------------------------------------------


@pytest.mark.parametrize("method", ["brute", "searchsorted"])
def test_find_indices_1d(rg, golden, method):
    xin, xout = cases.cases_find_indices()
    (r,) = rg.find_indices((xin,), (xout,), axis_input=-1, axis_output=-1, method=method)
    assert r.dtype == np.int64 and np.array_equal(r, golden[f"find/{method}"])
    (r,) = rg.find_indices((xin,), (xout,), axis_input=-1, axis_output=-1, method=method, fill_value=-1)
    assert np.array_equal(r, golden[f"find/{method}_fillm1"])
    # regridding/_find_indices/_tests/test_find_indices.py:54-91: sentinel regression
    x = np.linspace(0, 1, 5)
    p = np.array([-0.5, 0.0, 0.1, 0.9, 1.0, 1.5])
    (r,) = rg.find_indices((x,), (p,), method=method)
    big = np.iinfo(np.int64).max
    assert list(r) == [big, 0, 0, 3, 3, big]


def test_find_indices_2d_matches_reference_locators(rg, dev, oracle, golden):
    gi, _, _ = cases.case_2d("fam40")
    pts = golden["prim/points"]
    ri, rj = rg.find_indices(gi, (pts[:, 0].reshape(20, 20), pts[:, 1].reshape(20, 20)))
    assert np.array_equal(ri.reshape(-1), golden["prim/locate_brute"][:, 0])
    assert np.array_equal(rj.reshape(-1), golden["prim/locate_brute"][:, 1])
    assert np.array_equal(ri.reshape(-1), golden["prim/locate_secant"][:, 0])
    # larger random set against the oracle's brute locator, incl. points outside
    g = cases.curvilinear(70, 55, distort=0.01)
    rng = np.random.default_rng(9)
    P = 3000
    px = rng.uniform(g[0].min() - 0.1, g[0].max() + 0.1, P)
    py = rng.uniform(g[1].min() - 0.1, g[1].max() + 0.1, P)
    flat = rg.device.find_indices_2d(T(g[0], dev), T(g[1], dev), T(px, dev), T(py, dev), -1).cpu().numpy()
    big = np.iinfo(np.int64).max
    n_out = 0
    for k in range(P):
        i, j = oracle.index_of_point(g[0], g[1], px[k], py[k], "brute")
        want = -1 if i == big else i * (g[0].shape[1] - 1) + j
        assert flat[k] == want, (k, flat[k], want)
        n_out += want < 0
    assert 0 < n_out < P


def test_find_indices_2d_slow_pass_fallbacks(rg, dev, oracle, monkeypatch):
    """The paths config 5 never takes: the marks do not fit the queue (the slow pass scans the output for its
    sentinels), no raster-row index of the boundary (box walk), a `fill` equal to the first sentinel value, points on
    vertices / edges / outside next to the boundary."""
    g = cases.curvilinear(70, 55, distort=0.01)
    rng = np.random.default_rng(11)
    P = 4000
    px = rng.uniform(g[0].min() - 0.05, g[0].max() + 0.05, P)
    py = rng.uniform(g[1].min() - 0.05, g[1].max() + 0.05, P)
    # on vertices, on edge midpoints (tolerance zone of the cell walk) and non-finite
    k = rng.integers(0, 69, 300), rng.integers(0, 54, 300)
    px[:300], py[:300] = g[0][k], g[1][k]
    px[300:600] = 0.5 * (g[0][k] + g[0][k[0] + 1, k[1]])
    py[300:600] = 0.5 * (g[1][k] + g[1][k[0] + 1, k[1]])
    px[600], py[601], px[602] = np.nan, np.inf, -np.inf
    big = np.iinfo(np.int64).max
    want = np.empty(P, np.int64)
    for q in range(P):
        i, j = oracle.index_of_point(g[0], g[1], px[q], py[q], "brute")
        want[q] = -1 if i == big else i * (g[0].shape[1] - 1) + j
    args = (T(g[0], dev), T(g[1], dev), T(px, dev), T(py, dev))
    base = rg.device.find_indices_2d(*args, -1).cpu().numpy()
    assert np.array_equal(base, want)
    lo = np.iinfo(np.int64).min
    got = rg.device.find_indices_2d(*args, lo).cpu().numpy()
    assert np.array_equal(got, np.where(want < 0, lo, want))
    monkeypatch.setenv("RG_LOC_QUEUE_CAP", "7")
    assert np.array_equal(rg.device.find_indices_2d(*args, -1).cpu().numpy(), want)
    assert np.array_equal(rg.device.find_indices_2d(*args, lo + 1).cpu().numpy(), np.where(want < 0, lo + 1, want))
    monkeypatch.setenv("RG_LOC_NO_ROWMASK", "1")
    assert np.array_equal(rg.device.find_indices_2d(*args, -1).cpu().numpy(), want)
    monkeypatch.delenv("RG_LOC_QUEUE_CAP")
    assert np.array_equal(rg.device.find_indices_2d(*args, -1).cpu().numpy(), want)


def test_multilinear_1d(rg):
    """regridding/_weights/_weights_multilinear_test.py: linear functions are reproduced exactly."""
    x = np.linspace(0, 1, 11)
    p = np.linspace(0.05, 0.95, 7)
    vals = 3 * x + 1
    res = rg.regrid((x,), (p,), vals)
    assert np.allclose(res, 3 * p + 1)
    W = rg.weights((x,), (np.array([-0.2, 0.5, 1.3]),), bounds="nan")
    out = rg.regrid_from_weights(*W, vals)
    assert np.isnan(out[0]) and np.isnan(out[2]) and np.isclose(out[1], 2.5)
    with pytest.raises(ValueError):
        rg.weights((x,), (np.array([-0.2, 0.5]),), bounds="raise")
    ext = rg.regrid((x,), (np.array([-0.2, 1.3]),), vals)
    assert np.allclose(ext, 3 * np.array([-0.2, 1.3]) + 1)


def test_multilinear_1d_vs_reference_goldens(rg):
    """weights(method="multilinear") / regrid_from_weights / regrid against the reference's own output
    (tests/golden/golden_v3.npz, make_golden_v3.py): indices, weights (bit for bit, NaNs included) and applied values."""
    with np.load(cases.ROOT_GOLDEN / "golden_v3.npz") as z:
        G = {k: z[k] for k in z.files}
    for name, (x_in, x_out, w, bounds) in cases.cases_multilinear_1d().items():
        kw = dict(axis_input=-1, axis_output=-1) if x_in.ndim > 1 else {}
        W = rg.weights((x_in,), (x_out,), weights_input=w, method="multilinear", bounds=bounds, **kw)
        assert [*W[1], *W[2]] == list(G[f"ml1d/{name}/shapes"])
        flat = W[0].reshape(-1)
        for d in range(flat.size):
            ii, io, v = flat[d]
            assert ii.dtype == np.int64 and io.dtype == np.int64 and v.dtype == np.float64
            assert np.array_equal(ii, G[f"ml1d/{name}/{d}/ii"]), (name, d)
            assert np.array_equal(io, G[f"ml1d/{name}/{d}/io"]), (name, d)
            assert np.array_equal(v, G[f"ml1d/{name}/{d}/v"], equal_nan=True), (name, d)
        vals = np.random.default_rng(3).random(W[1])
        assert np.array_equal(rg.regrid_from_weights(*W, vals, **kw), G[f"ml1d/{name}/apply"], equal_nan=True), name
        got = rg.regrid((x_in,), (x_out,), vals, method="multilinear", bounds=bounds, **kw)
        assert np.array_equal(got, G[f"ml1d/{name}/regrid"], equal_nan=True), name
    x = np.linspace(0, 1, 11)
    with pytest.raises(ValueError, match="2 of the output points fall outside"):
        rg.weights((x,), (np.array([-0.2, 0.5, 1.3]),), bounds="raise")


def test_build2d_batched_equals_single_builds(rg, dev):
    """rg_build2d_batched (per-slice grids enqueued back to back, no host sync; BASELINE config 4's inner loop) gives
    the same triplets as one rg_build2d_* build per slice; also through the single-process form of the sharded entry."""
    from regridding_b200 import _parallel

    slices = []
    for f in range(5):
        gi, go = cases.benchmark_family(97, 90, 101, distorted=True, angle=0.4 + 0.01 * f, phase=float(f))
        co = cases.perturb_like_reference(go, (-1, -2), 42)
        slices.append((gi[0], gi[1], co[0], co[1]))
    want = [rg.device.build_weights_2d(*sl, device=dev) for sl in slices]
    for attempt in range(2):   # first call: estimated buffer sizes; second: learned ones
        got = rg.device.build_weights_2d_batched([tuple(T(a, dev) for a in sl) for sl in slices], device=dev)
        for a, b in zip(got, want):
            assert torch.equal(a.indices_input, b.indices_input) and torch.equal(a.indices_output, b.indices_output)
            assert torch.equal(a.values, b.values)
    sharded = _parallel.build_weights_2d_slices(slices, device=dev, chunk=2)
    assert sorted(sharded) == list(range(5))
    for k, b in enumerate(want):
        assert torch.equal(sharded[k].values, b.values) and torch.equal(sharded[k].indices_output, b.indices_output)


def test_ndarray_linear_interpolation_vs_reference_goldens(rg):
    """regridding.ndarray_linear_interpolation (regridding/_interp_ndarray.py:11-297) on the device against the
    reference's own output, bit for bit (1D plain IEEE; 2D the fastmath contraction the reference's JIT emits), for
    every axis / axis_indices combination the reference's tests use, incl. extrapolation; plus its error behaviour."""
    with np.load(cases.ROOT_GOLDEN / "golden_v3.npz") as z:
        G = {k: z[k] for k in z.files}
    for name, (a, indices, kw) in cases.cases_interp_ndarray().items():
        got = rg.ndarray_linear_interpolation(a, indices, **kw)
        want = G[f"interp/{name}"]
        assert got.shape == want.shape and got.dtype == np.float64, name
        assert np.array_equal(got, want), (name, float(np.max(np.abs(got - want))))
    with pytest.raises(ValueError, match="must match the number of elements in axis"):
        rg.ndarray_linear_interpolation(np.zeros((3, 4)), (np.zeros(2),))
    with pytest.raises(NotImplementedError):
        rg.ndarray_linear_interpolation(np.zeros((3, 4, 5)), (np.zeros(2), np.zeros(2), np.zeros(2)))
    # regridding/_tests/test_interp_ndarray.py:12-88: agrees with scipy.ndimage.map_coordinates inside the array
    scipy_ndimage = pytest.importorskip("scipy.ndimage")
    a = np.random.default_rng(0).random((10, 11))
    x, y = np.broadcast_arrays(np.linspace(0, 9, 100)[:, None], np.linspace(0, 10, 5)[None, :])
    got = rg.ndarray_linear_interpolation(a, (x, y), axis=(0, 1))
    assert np.allclose(got, scipy_ndimage.map_coordinates(a, np.stack([x, y]), order=1))


def test_weights_1d_many_spectra_are_built_in_chunks(rg, oracle, monkeypatch):
    """Stacked 1D conservative builds run in chunks bounded by a byte budget: same triplets as the oracle."""
    from regridding_b200 import _weights

    xin, xout, _ = cases.cases_1d()["spectra"]
    monkeypatch.setattr(_weights, "_CHUNK_BYTES_1D", 24 * (xin.shape[1] + xout.shape[1]) * 7)  # 7 spectra per chunk
    W = rg.weights((xin,), (xout,), axis_input=-1, axis_output=-1, method="conservative")
    ref = oracle.weights_conservative_1d_batched(xin, xout)
    for s in range(xin.shape[0]):
        for a, b in zip(W[0][s], ref[s]):
            assert np.array_equal(a, b)
    vals = np.random.default_rng(0).random(W[1])
    out = rg.regrid_from_weights(*W, vals, axis_input=-1, axis_output=-1)
    assert np.array_equal(out, rg.regrid((xin,), (xout,), vals, axis_input=-1, axis_output=-1, method="conservative"))


# ---------------------------------------------------------------------------
# transposed weights (regridding/_weights/_weights_transposed)
# ---------------------------------------------------------------------------


@pytest.mark.parametrize("name", ["fam40", "winput", "coarsen"])
def test_transposed_weights_2d_vs_reference(rg, oracle, golden, name):
    gi, go, w = cases.case_2d(name)
    W = rg.weights(gi, go, weights_input=w, method="conservative")
    vals = np.random.default_rng(0).random((3, *W[1]))
    fwd = rg.regrid_from_weights(*W, vals)
    # plain transpose: index arrays swapped, values shared; the apply must NOT pick up the forward matrix
    Wt = rg.transpose_weights(W)
    assert Wt[1] == W[2] and Wt[2] == W[1] and Wt[0][()][2] is W[0][()][2] and Wt[0][()][0] is W[0][()][1]
    assert cases.sha(*Wt[0][()]) == str(golden[f"t2d/{name}/plain_sha"])
    assert np.array_equal(rg.regrid_from_weights(*Wt, fwd), golden[f"t2d/{name}/plain_apply"])
    # conservative transpose: the reference's values, bit for bit
    Wc = rg.transpose_weights_conservative(W, gi, go, weights_input=w)
    ii, io, v = Wc[0][()]
    assert np.array_equal(ii, golden[f"t2d/{name}/cons_ii"]) and np.array_equal(io, golden[f"t2d/{name}/cons_io"])
    assert np.array_equal(v, golden[f"t2d/{name}/cons_v"])
    assert [*Wc[1], *Wc[2]] == list(golden[f"t2d/{name}/cons_shapes"])
    assert np.array_equal(rg.regrid_from_weights(*Wc, fwd), golden[f"t2d/{name}/cons_apply"])
    # oracle restatement (NumPy arithmetic + the C grid_volume)
    fi, fo, fv = W[0][()]
    ref = oracle.transpose_weights_conservative_values(fi, fo, fv, oracle.grid_volume(*gi).reshape(-1),
                                                       oracle.grid_volume(*go).reshape(-1),
                                                       None if w is None else np.asarray(w, dtype=float).reshape(-1))
    assert np.array_equal(v, ref)


def test_transposed_weights_batched_and_1d_vs_reference(rg, golden):
    gi, go = cases.case_2d_batched()
    kw = dict(axis_input=(1, 2), axis_output=(1, 2))
    W = rg.weights(gi, go, method="conservative", **kw)
    Wc = rg.transpose_weights_conservative(W, gi, go, **kw)
    for f in range(3):
        assert np.array_equal(Wc[0][f][2], golden[f"t2d_batched/{f}/v"])
        assert cases.sha(*Wc[0][f]) == str(golden[f"t2d_batched/{f}/sha"])
    vals = np.random.default_rng(0).random(W[1])
    fwd = rg.regrid_from_weights(*W, vals, **kw)
    assert np.array_equal(rg.regrid_from_weights(*Wc, fwd, **kw), golden["t2d_batched/apply"])
    for name in ("spectra", "spectra_w", "descending_nonuniform", "descending_both"):
        xin, xout, w = cases.cases_1d()[name]
        k1 = dict(axis_input=-1, axis_output=-1)
        W1 = rg.weights((xin,), (xout,), weights_input=w, method="conservative", **k1)
        W1c = rg.transpose_weights_conservative(W1, (xin,), (xout,), weights_input=w, **k1)
        flat = W1c[0].reshape(-1)
        assert np.array_equal(np.concatenate([e[2] for e in flat]), golden[f"t1d/{name}/v"]), name
        assert np.array_equal(np.concatenate([e[0] for e in flat]), golden[f"t1d/{name}/ii"])
        vals1 = np.random.default_rng(0).random(W1[1])
        fwd1 = rg.regrid_from_weights(*W1, vals1, **k1)
        assert np.array_equal(rg.regrid_from_weights(*W1c, fwd1, **k1), golden[f"t1d/{name}/apply"], equal_nan=True), name


def test_find_indices_2d_config5_shape_properties(rg, dev):
    """BASELINE.json config 5 at reduced size (1024^2 vertices, 2048^2 points over the full bounding box):
    every located point lies inside the cell it was assigned to, and every point reported outside lies
    outside the grid (its winding number around the boundary polygon is 0)."""
    n, m = 1024, 2048
    gi, _ = cases.benchmark_family(n, distorted=True)
    X, Y = gi
    px = np.broadcast_to(np.linspace(X.min(), X.max(), m)[:, None], (m, m)).copy()
    py = np.broadcast_to(np.linspace(Y.min(), Y.max(), m)[None, :], (m, m)).copy()
    flat = rg.device.find_indices_2d(T(X, dev), T(Y, dev), T(px, dev), T(py, dev), -1).cpu().numpy()
    inside = flat >= 0
    assert 0.5 < inside.mean() < 0.65  # the rotated grid covers ~58 % of its bounding box
    ci, cj = flat[inside] // (n - 1), flat[inside] % (n - 1)
    qx, qy = px[inside], py[inside]
    # counter-clockwise (in index space) corners of the assigned cells
    cx = np.stack([X[ci, cj], X[ci + 1, cj], X[ci + 1, cj + 1], X[ci, cj + 1]])
    cy = np.stack([Y[ci, cj], Y[ci + 1, cj], Y[ci + 1, cj + 1], Y[ci, cj + 1]])
    cross = [(cx[(k + 1) % 4] - cx[k]) * (qy - cy[k]) - (cy[(k + 1) % 4] - cy[k]) * (qx - cx[k]) for k in range(4)]
    cross = np.stack(cross)
    sign = np.sign(cross.sum(axis=0))  # orientation of the cell
    assert np.all(cross * sign >= -1e-12), "a point was assigned to a cell that does not contain it"
    # outside points: winding number around the boundary polygon
    bx = np.concatenate([X[:-1, 0], X[-1, :-1], X[:0:-1, -1], X[0, :0:-1]])
    by = np.concatenate([Y[:-1, 0], Y[-1, :-1], Y[:0:-1, -1], Y[0, :0:-1]])
    ox, oy = px[~inside][::97], py[~inside][::97]
    ang = np.zeros(ox.shape)
    for k in range(bx.size):
        x0, y0 = bx[k] - ox, by[k] - oy
        x1, y1 = bx[(k + 1) % bx.size] - ox, by[(k + 1) % bx.size] - oy
        ang += np.arctan2(x0 * y1 - x1 * y0, x0 * x1 + y0 * y1)
    assert np.all(np.abs(ang) < 1.0), "a point inside the grid was reported as outside"


def test_apply_degenerate_sizes(rg, dev):
    """Odd (not 16-byte friendly) grid sizes are staged too -- odd frames one cell later, the unpaired last frame
    through the generic kernel -- misaligned value pointers take the generic path; zero frames are a no-op."""
    gi = cases.curvilinear(12, 10)                     # 11 x 9 = 99 cells: odd input width and odd n_in
    go = cases.rectilinear_over(gi[0], gi[1], 8, 8)    # 7 x 7 cells: odd output width
    dw = rg.device.build_weights_2d(gi[0], gi[1], go[0], go[1], device=dev)
    plan = dw.plan((11, 9), (7, 7))
    assert plan.n_generic_tiles == 0  # no odd-width cliff: the bulk-copy kernel stages flat, even-aligned spans
    for F in (1, 2, 5, 8, 9, 16, 23):
        x = torch.rand((F, dw.n_in), dtype=torch.float64, device=dev)
        assert torch.equal(rg.device.apply_planned(plan, x), rg.device.apply_csr(dw.csr(), x)), F
    # values that start 8 bytes off a 16-byte boundary
    big = torch.rand((7, dw.n_in), dtype=torch.float64, device=dev)
    x = big[1:]
    assert x.data_ptr() % 16 == 8 and x.is_contiguous()
    assert torch.equal(rg.device.apply_planned(plan, x), rg.device.apply_csr(dw.csr(), x))
    # output rows that start 8 bytes off (odd w_out and a misaligned out buffer)
    outbuf = torch.empty((7, dw.n_out), dtype=torch.float64, device=dev)
    x = torch.rand((6, dw.n_in), dtype=torch.float64, device=dev)
    got = rg.device.apply_planned(plan, x, out=outbuf[1:])
    assert torch.equal(got, rg.device.apply_csr(dw.csr(), x))
    empty = torch.empty((0, dw.n_in), dtype=torch.float64, device=dev)
    assert rg.device.apply_planned(plan, empty).shape == (0, dw.n_out)
    assert rg.device.apply_csr(dw.csr(), empty).shape == (0, dw.n_out)


@pytest.mark.parametrize("shape_in,shape_out", [((100, 100), (120, 80)), ((101, 98), (57, 131)), ((64, 333), (200, 31))])
def test_apply_bulk_odd_and_ragged_shapes(rg, dev, oracle, shape_in, shape_out):
    """Vertex counts that make n_in / w_in / w_out odd in every combination, with partial tiles on both output axes:
    staged apply == generic apply == oracle, for frame counts around the 8-frame stage and the 512-frame block."""
    gi = cases.curvilinear(*shape_in, distort=0.01)
    go = cases.rectilinear_over(gi[0], gi[1], *shape_out)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    dw = rg.device.build_weights_2d(gi[0], gi[1], co[0], co[1], device=dev)
    cin = (shape_in[0] - 1, shape_in[1] - 1)
    cout = (shape_out[0] - 1, shape_out[1] - 1)
    plan = dw.plan(cin, cout)
    if shape_in[1] < 300:  # (the strongly anisotropic pair has footprints beyond a staged tile: generic kernel there)
        assert plan.n_generic_tiles <= plan.n_tiles // 4
    for F in (3, 8, 15, 513):
        x = torch.rand((F, dw.n_in), dtype=torch.float64, device=dev)
        a = rg.device.apply_planned(plan, x)
        assert torch.equal(a, rg.device.apply_csr(dw.csr(), x)), F
    ii, io, v = dw.to_host()
    vals = np.random.default_rng(1).random((9, dw.n_in))
    ref = oracle.regrid_from_weights(ii, io, v, vals, dw.n_out)
    assert np.array_equal(rg.device.apply_planned(plan, T(vals, dev)).cpu().numpy(), ref)
