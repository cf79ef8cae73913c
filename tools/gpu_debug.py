"""Verbose CUDA-vs-oracle/golden comparison for development (run on the GPU box)."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import regridding_b200 as rg
from regridding_b200 import _device
from oracle import oracle
from tests import cases

G = dict(np.load(ROOT / "tests/golden/golden_v1.npz"))
dev = torch.device("cuda", 0)
print(torch.cuda.get_device_name(0), flush=True)

def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"   {name}: SHAPE {a.shape} vs {b.shape}"); return False
    eq = np.array_equal(a, b)
    if not eq:
        bad = np.flatnonzero(a.reshape(-1) != b.reshape(-1))
        print(f"   {name}: {bad.size}/{a.size} differ; first {bad[:5]} got {a.reshape(-1)[bad[:5]]} want {b.reshape(-1)[bad[:5]]}")
        if a.dtype.kind == "f":
            print(f"      max abs {np.abs(a-b).max():.3e}")
    return eq

# area
for name in ("fam40", "coarsen", "flipx"):
    g, _, _ = cases.case_2d(name)
    a = _device.grid_area(*g, device=dev).cpu().numpy()
    print("area", name, cmp("area", a, G[f"prim/volume_{name}"]))

for name in cases.CASES_2D:
    gi, go, w = cases.case_2d(name)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    t = time.time()
    try:
        dw = _device.build_weights_2d(gi[0], gi[1], co[0], co[1], w, device=dev)
    except Exception as e:
        print(name, "BUILD FAILED", repr(e)); continue
    torch.cuda.synchronize(); dt = time.time() - t
    ii, io, v = dw.to_host()
    print(name, "nnz", v.size, "golden", int(G[f"c2d/{name}/nnz"]), "frag", dw.stats, f"{dt*1e3:.1f} ms")
    raw = oracle.weights_conservative_2d(gi, co, w)
    oi, oo, ov = oracle.coalesce(*raw)
    ok = cmp("ii", ii, oi) & cmp("io", io, oo) & cmp("v", v, ov)
    print("   sha final", cases.sha(ii, io, v) == str(G[f"c2d/{name}/final_sha"]), "raw n", raw[2].size)
    if ok:
        vals = np.random.default_rng(0).random((3, dw.n_in))
        out = _device.apply_csr(dw.csr(), torch.from_numpy(vals).to(dev)).cpu().numpy()
        ref = oracle.regrid_from_weights(oi, oo, ov, vals, dw.n_out)
        cmp("apply", out, ref)
        shape_out = tuple(G[f"c2d/{name}/shape_out"])
        print("   apply sha", cases.sha(out.reshape(3, *shape_out)) == str(G[f"c2d/{name}/apply_sha"]))

# public API
gi, go, w = cases.case_2d("fam40")
W = rg.weights(gi, go, method="conservative")
print("api weights fam40", cases.sha(*W[0][()]) == str(G["c2d/fam40/final_sha"]), W[1], W[2])
vals = np.random.default_rng(0).random((3, *W[1]))
res = rg.regrid_from_weights(*W, vals)
print("api apply", cmp("apply", res, G["c2d/fam40/apply"]))
res2 = rg.regrid(gi, go, vals, method="conservative")
print("api regrid", cmp("regrid", res2, G["c2d/fam40/apply"]))

# 1D
for name, (xin, xout, w) in cases.cases_1d().items():
    W = rg.weights((xin,), (xout,), axis_input=-1, axis_output=-1, weights_input=w, method="conservative")
    flat = W[0].reshape(-1)
    ok = cmp("ii", np.concatenate([e[0] for e in flat]), G[f"c1d/{name}/ii"]) & \
         cmp("io", np.concatenate([e[1] for e in flat]), G[f"c1d/{name}/io"]) & \
         cmp("v", np.concatenate([e[2] for e in flat]), G[f"c1d/{name}/v"])
    vals = np.random.default_rng(0).random(W[1])
    res = rg.regrid_from_weights(*W, vals, axis_input=-1, axis_output=-1)
    ok &= cmp("apply", res, G[f"c1d/{name}/apply"])
    fused = _device.regrid1d_conservative(torch.from_numpy(xin).to(dev), torch.from_numpy(xout).to(dev),
                                          torch.from_numpy(vals).to(dev),
                                          None if w is None else torch.from_numpy(w).to(dev)).cpu().numpy()
    ok &= cmp("fused", fused, G[f"c1d/{name}/apply"])
    print("1d", name, ok)

# find_indices
xin, xout = cases.cases_find_indices()
for method in ("brute", "searchsorted"):
    (r,) = rg.find_indices((xin,), (xout,), axis_input=-1, axis_output=-1, method=method)
    print("find", method, cmp("idx", r, G[f"find/{method}"]))
    (r,) = rg.find_indices((xin,), (xout,), axis_input=-1, axis_output=-1, method=method, fill_value=-1)
    print("find fill", method, cmp("idx", r, G[f"find/{method}_fillm1"]))
gi, _, _ = cases.case_2d("fam40")
pts = G["prim/points"]
ri, rj = rg.find_indices(gi, (pts[:, 0].reshape(20, 20), pts[:, 1].reshape(20, 20)))
print("find2d", cmp("i", ri.reshape(-1), G["prim/locate_brute"][:, 0]), cmp("j", rj.reshape(-1), G["prim/locate_brute"][:, 1]))

# timing at scale
for n in (513, 1025, 2049):
    (gi, go) = cases.benchmark_family(n, distorted=True)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    xi, yi, xo, yo = (torch.from_numpy(a).to(dev) for a in (*gi, *co))
    for rep in range(3):
        torch.cuda.synchronize(); t = time.time()
        dw = _device.build_weights_2d(xi, yi, xo, yo, device=dev)
        torch.cuda.synchronize(); dt = time.time() - t
    print(f"build {n}: {dt*1e3:.2f} ms  {(n-1)**2/dt/1e6:.1f} Mcells/s", dw.stats, flush=True)
    csr = dw.csr()
    F = 64
    vin = torch.rand((F, dw.n_in), dtype=torch.float64, device=dev)
    out = torch.empty((F, dw.n_out), dtype=torch.float64, device=dev)
    for rep in range(3):
        torch.cuda.synchronize(); t = time.time()
        _device.apply_csr(csr, vin, out)
        torch.cuda.synchronize(); dt = time.time() - t
    byt = 8 * F * (dw.n_in + dw.n_out) + 12 * dw.nnz + 4 * (dw.n_out + 1)
    print(f"apply {n} F={F}: {dt*1e3:.2f} ms  {byt/dt/1e9:.0f} GB/s")
    plan = dw.plan((n - 1, n - 1), (n - 1, n - 1))
    out2 = torch.empty_like(out)
    for F2 in (64, 256):
        vin2 = torch.rand((F2, dw.n_in), dtype=torch.float64, device=dev)
        out2 = torch.empty((F2, dw.n_out), dtype=torch.float64, device=dev)
        for rep in range(3):
            torch.cuda.synchronize(); t = time.time()
            _device.apply_planned(plan, vin2, out2)
            torch.cuda.synchronize(); dt = time.time() - t
        byt = 8 * F2 * (dw.n_in + dw.n_out) + 12 * dw.nnz + 4 * (dw.n_out + 1)
        ref2 = _device.apply_csr(csr, vin2)
        print(f"planned {n} F={F2}: {dt*1e3:.2f} ms  {byt/dt/1e9:.0f} GB/s  generic tiles {plan.n_generic_tiles}/{plan.n_tiles} equal {torch.equal(ref2, out2)}")
