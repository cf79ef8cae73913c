"""CPU tests of the host-side logic that mirrors the reference's Python layer (no GPU needed)."""

import numpy as np
import pytest

from regridding_b200 import _cache, _parallel, _util
from tests import cases


def test_perturbation_matches_reference_stream(golden):
    """The seeded 1e-9 jitter (regridding/_util.py:121-129) -- golden SHA of the reference's output."""
    for name in cases.CASES_2D:
        gi, go, _ = cases.case_2d(name)
        _, co, axis_in, axis_out, *_ = _util.normalize_input_output_coordinates(gi, go, perturb=True, seed=42)
        assert axis_in == (-1, -2) and axis_out == (-1, -2)
        assert cases.sha(*(np.ascontiguousarray(c) for c in co)) == str(golden[f"c2d/{name}/perturbed_sha"])


def test_perturbation_stream_order_batched():
    """x for ALL orthogonal slices first, then y; per-slice spread."""
    gi, go = cases.case_2d_batched()
    _, co, *_ = _util.normalize_input_output_coordinates(gi, go, axis_input=(1, 2), axis_output=(1, 2),
                                                         perturb=True, seed=42)
    ref = cases.perturb_like_reference(go, (-1, -2), 42)
    assert np.array_equal(co[0], ref[0]) and np.array_equal(co[1], ref[1])
    rng = np.random.default_rng(42)
    zx = rng.standard_normal(go[0].shape)
    spread = np.ptp(go[0], axis=(-1, -2), keepdims=True)
    assert np.array_equal(co[0], go[0] + (spread * 1e-9) * zx)
    # a slice built alone gets a different jitter than inside the batch
    _, alone, *_ = _util.normalize_input_output_coordinates((gi[0][1], gi[1][1]), (go[0][1], go[1][1]),
                                                            perturb=True, seed=42)
    assert not np.array_equal(alone[0], co[0][1])


def test_normalize_shapes_and_axes():
    x = np.zeros((7, 5, 6))
    y = np.zeros((7, 5, 6))
    xo = np.zeros((1, 4, 3))
    r = _util.normalize_input_output_coordinates((x, y), (xo, xo), axis_input=(1, 2), axis_output=(-2, -1))
    ci, co, ai, ao, si, so, orth = r
    assert ai == (-1, -2) and ao == (-1, -2) and orth == (7,)
    assert ci[0].shape == (7, 5, 6) and co[0].shape == (7, 4, 3)
    # resampled axis in the middle
    r = _util.normalize_input_output_coordinates((np.zeros((3, 9, 2)),), (np.zeros((3, 4, 2)),), axis_input=1,
                                                 axis_output=1)
    assert r[2] == (-2,) and r[6] == (3, 2) and r[0][0].shape == (3, 9, 2) and r[1][0].shape == (3, 4, 2)
    # bare arrays are wrapped
    r = _util.normalize_input_output_coordinates(np.zeros(5), np.zeros(3))
    assert r[2] == (-1,) and r[6] == ()
    assert _util.normalize_axis(None, 3) == (-3, -2, -1)
    assert _util.normalize_axis((0, -1), 3) == (-3, -1)
    assert _util._embed((7, 2), (-1, -3), {-1: 5, -3: 9}) == (7, 9, 2, 5)


def test_argument_errors_like_reference():
    """regridding/_util.py:68-84 and _weights.py:184 / _find_indices.py:125 -- raised before any GPU work."""
    x = np.zeros((4, 4))
    with pytest.raises(ValueError, match="number of axes"):
        _util.normalize_input_output_coordinates((x, x), (x, x), axis_input=(0, 1), axis_output=(0,))
    with pytest.raises(ValueError, match="coordinates_input"):
        _util.normalize_input_output_coordinates((x,), (x,), axis_input=(0, 1), axis_output=(0, 1))
    with pytest.raises(ValueError, match="coordinates_output"):
        _util.normalize_input_output_coordinates((x,), (x, x), axis_input=(0,), axis_output=(0,))
    import regridding_b200 as rg

    with pytest.raises(ValueError, match="unrecognized method"):
        rg.weights((x, x), (x, x), method="bogus")
    with pytest.raises(ValueError, match="not recognized"):
        rg.find_indices((np.zeros(3),), (np.zeros(3),), method="bogus")
    with pytest.raises(ValueError, match="bounds"):
        rg.weights((np.zeros(3),), (np.zeros(3),), bounds="bogus")


def test_no_cpu_fallback_without_cuda():
    import torch

    import regridding_b200 as rg

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    x = np.linspace(0, 1, 5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rg.weights((x,), (x,), method="conservative")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rg.find_indices((x,), (x,))


def test_cache_is_identity_keyed():
    ii, io, v = np.arange(3), np.arange(3) + 1, np.arange(3.0)
    token = object()
    _cache.remember((ii, io, v), token)
    assert _cache.lookup((ii, io, v)) is token
    assert _cache.lookup((ii, io, v.copy())) is None
    # transpose_weights re-uses the values array with the index arrays swapped: must not hit the forward matrix
    assert _cache.lookup((io, ii, v)) is None
    assert _cache.lookup((ii.copy(), io, v)) is None
    # the transposed use gets its own entry and does not evict the forward one
    token_t = object()
    _cache.remember((io, ii, v), token_t)
    assert _cache.lookup((io, ii, v)) is token_t and _cache.lookup((ii, io, v)) is token
    # an in-place edit after the upload invalidates the entry (arrays handed out by weights() are read-only on top)
    v[1] = 7.0
    assert _cache.lookup((ii, io, v)) is None
    _cache.freeze((ii, io, v))
    with pytest.raises(ValueError):
        v[0] = 1.0
    n = len(_cache._entries)
    del v
    import gc

    gc.collect()
    assert len(_cache._entries) == n - 1


def test_cache_is_bounded_by_device_bytes(monkeypatch):
    class Fake:
        device = None

        def __init__(self, nbytes):
            self.nbytes = nbytes

        def device_bytes(self):
            return self.nbytes

    _cache.clear()
    monkeypatch.setenv("REGRID_B200_CACHE_BYTES", "1000")
    keep = []
    for k in range(4):
        el = (np.arange(3) + k, np.arange(3), np.arange(3.0) + k)
        keep.append(el)
        _cache.remember(el, Fake(400))
    assert _cache.lookup(keep[0]) is None and _cache.lookup(keep[1]) is None  # evicted, oldest first
    assert _cache.lookup(keep[2]) is not None and _cache.lookup(keep[3]) is not None
    _cache.clear()


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 64, 1000, 2048):
        for w in (1, 2, 3, 8):
            parts = [_parallel.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert _parallel.band_cells(2048, 2048, 3, 8) == (3 * 256 * 2048, 4 * 256 * 2048)


def test_packed_weights_roundtrip(tmp_path):
    """PackedWeights <-> reference layout (object ndarray of tuples) and the memory-mapped file format."""
    from regridding_b200._packed import PackedWeights

    rng = np.random.default_rng(0)
    shape_orth = (2, 3)
    w = np.empty(shape_orth, dtype=object)
    for idx in np.ndindex(*shape_orth):
        k = int(rng.integers(0, 6))  # some slices are empty
        w[idx] = (rng.integers(-4, 9, k), rng.integers(0, 7, k), rng.random(k))
    ref = (w, (2, 3, 9), (2, 3, 7))
    p = PackedWeights.from_reference(ref)
    assert len(p) == 6 and p.nnz == sum(len(e[2]) for e in w.reshape(-1))
    back, s_in, s_out = p.to_reference()
    assert back.shape == shape_orth and s_in == (2, 3, 9) and s_out == (2, 3, 7)
    for idx in np.ndindex(*shape_orth):
        for a, b in zip(back[idx], w[idx]):
            assert a.dtype == b.dtype and np.array_equal(a, b)
    path = tmp_path / "weights.rgpw"
    p.save(path)
    for mmap in (True, False):
        q = PackedWeights.load(path, mmap=mmap)
        assert q.shape_input == p.shape_input and q.shape_output == p.shape_output and q.shape_orthogonal == shape_orth
        for name in ("offsets", "indices_input", "indices_output", "values"):
            assert np.array_equal(getattr(q, name), getattr(p, name))
    # an all-empty weights object and a corrupted file
    e = PackedWeights.from_reference((np.array((np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0)), dtype=object)[None][:0].reshape(0), (0, 4), (0, 5)))
    assert len(e) == 0 and e.nnz == 0
    path.write_bytes(b"not a weights file")
    with pytest.raises(ValueError):
        PackedWeights.load(path)
    with pytest.raises(ValueError):
        PackedWeights(np.array([0, 2]), np.zeros(1), np.zeros(1), np.zeros(1), (3,), (3,), ())


def test_band_bounds_partition_the_input_cells():
    """Input-row bands of the sharded build: contiguous, ordered, covering every cell once, empty bands allowed."""
    for ncx, ncy, w in [(2048, 2048, 8), (99, 100, 5), (3, 7, 8), (1, 1, 2)]:
        b = _parallel.band_bounds(ncx, ncy, w)
        assert len(b) == w + 1 and b[0] == 0 and b[-1] == ncx * ncy
        assert all(b[r] <= b[r + 1] and b[r] % ncy == 0 for r in range(w))
        assert [b[r + 1] - b[r] for r in range(w)] == [
            (_parallel.shard_range(ncx, r, w)[1] - _parallel.shard_range(ncx, r, w)[0]) * ncy for r in range(w)]
        for r in range(w):
            assert _parallel.band_cells(ncx, ncy, r, w) == (b[r], b[r + 1])


def test_sharded_build_rejects_unknown_exchange(monkeypatch):
    import torch.distributed as dist

    monkeypatch.setattr(_parallel, "world", lambda group=None: (0, 2))
    with pytest.raises(ValueError, match="exchange"):
        _parallel.build_weights_2d_sharded(np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3)),
                                           exchange="carrier-pigeon")
    del dist
