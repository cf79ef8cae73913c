"""Times the BASELINE.json configurations other than the bench's config 3 (development / DESIGN.md numbers).
python tools/run_configs.py [out.json]"""
import json, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import regridding_b200 as rg
from regridding_b200 import _device
from tests import cases

dev = torch.device("cuda", 0)
out = {}

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

# ---- config 1: 100x100 vertices -> 120x80, one image
gi, go, _ = cases.case_2d("fam100")
co = cases.perturb_like_reference(go, (-1, -2), 42)
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (*gi, *co)]
ms_b, dw = timed(lambda: _device.build_weights_2d(*t, device=dev))
plan = dw.plan((99, 99), (119, 79))
x = torch.rand((1, dw.n_in), dtype=torch.float64, device=dev)
ms_a, _ = timed(lambda: _device.apply_planned(plan, x), reps=20)
t0 = time.perf_counter(); rg.regrid(gi, go, np.random.default_rng(0).random((99, 99)), method="conservative"); t_api = time.perf_counter() - t0
out["config1"] = {"build_ms": ms_b, "apply_ms_1_frame": ms_a, "nnz": dw.nnz, "regrid_api_s_host_to_host": t_api}

# ---- config 2: S spectra x 4096 bins, per-spectrum grids (fused 1D conservative regrid)
S, n = 65536, 4097
g = torch.Generator(device=dev); g.manual_seed(0)
base = torch.linspace(4000.0, 7000.0, n, dtype=torch.float64, device=dev)
xin = base * (1 + 1e-4 * torch.randn((S, 1), dtype=torch.float64, device=dev, generator=g)) + 0.3 * torch.sin(base / 500 + torch.rand((S, 1), dtype=torch.float64, device=dev, generator=g))
xout = torch.linspace(4001.0, 6999.0, n, dtype=torch.float64, device=dev) + 0.05 * torch.rand((S, 1), dtype=torch.float64, device=dev, generator=g)
vals = torch.rand((S, n - 1), dtype=torch.float64, device=dev, generator=g)
res = torch.empty((S, n - 1), dtype=torch.float64, device=dev)
ms, _ = timed(lambda: _device.regrid1d_conservative(xin, xout, vals, out=res), reps=10)
byt = S * (8 * n + 8 * n + 8 * (n - 1) + 8 * (n - 1))
flux_in = (vals * 1.0).sum(dim=1)
out["config2"] = {"spectra": S, "bins": n - 1, "fused_regrid_ms": ms, "spectra_per_s": S / ms * 1e3, "algorithmic_GBps": byt / ms / 1e6,
                  "frac_of_hbm_peak": byt / ms / 1e6 / 6549.8, "full_1M_spectra_s_extrapolated": ms * (1e6 / S) / 1e3}
ms_w, w = timed(lambda: _device.cons1d_batched(xin[:8192], xout[:8192]), reps=3)
out["config2"]["weights_materialised_ms_per_8192_spectra"] = ms_w
del xin, xout, vals, res, w
torch.cuda.empty_cache()

# ---- config 4: frames with their own curvilinear grids (per-slice builds)
n4, F4 = 2049, 4
builds = []
for f in range(F4):
    gi4, go4 = cases.benchmark_family(n4, distorted=True, angle=0.4 + 0.002 * f, phase=float(f))
    co4 = cases.perturb_like_reference(go4, (-1, -2), 42)
    tt = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (*gi4, *co4)]
    ms, dw4 = timed(lambda: _device.build_weights_2d(*tt, device=dev), reps=2, warm=1)
    builds.append(ms)
out["config4"] = {"frames_timed": F4, "build_ms_per_frame": builds, "64_frames_on_1_gpu_s": float(np.mean(builds)) * 64 / 1e3,
                  "64_frames_on_8_gpus_s": float(np.mean(builds)) * 8 / 1e3, "note": "frames shard across ranks with no collective; the host jitter stream (0.14 s per frame) is serial by the reference's definition"}
del dw4, tt
torch.cuda.empty_cache()

# ---- config 5: cell location of 8192^2 points in a 4096^2-vertex curvilinear grid
gi5, _ = cases.benchmark_family(4096, distorted=True)
X, Y = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in gi5)
m = 8192
for name, scale in (("inside_0.7_bbox", 0.7), ("full_bbox", 1.0)):
    cx, cy = float(X.min() + X.max()) / 2, float(Y.min() + Y.max()) / 2
    hx, hy = float(X.max() - X.min()) / 2 * scale, float(Y.max() - Y.min()) / 2 * scale
    px = torch.linspace(cx - hx, cx + hx, m, dtype=torch.float64, device=dev)[:, None].expand(m, m).contiguous()
    py = torch.linspace(cy - hy, cy + hy, m, dtype=torch.float64, device=dev)[None, :].expand(m, m).contiguous()
    ms, idx = timed(lambda: _device.find_indices_2d(X, Y, px, py, -1), reps=2, warm=1)
    inside = float((idx >= 0).double().mean())
    out["config5_" + name] = {"points": m * m, "ms": ms, "Mpoints_per_s": m * m / ms / 1e3, "fraction_inside": inside,
                              "algorithmic_GBps": (16 + 8) * m * m / ms / 1e6}
    del idx
    # ... + bilinear weights (4 per point) + the four-point gather of one frame of vertex values
    vals5 = torch.rand((1, 4096 * 4096), dtype=torch.float64, device=dev)
    pxf, pyf = px.reshape(-1), py.reshape(-1)

    def ml():
        idx4, w4, _ = _device.multilinear2d_weights(X, Y, pxf, pyf, "nan")
        return _device.ell4_apply(idx4, w4, vals5)
    ms_ml, res = timed(ml, reps=2, warm=1)
    idx4, w4, _ = _device.multilinear2d_weights(X, Y, pxf, pyf, "nan")
    ms_ap, _ = timed(lambda: _device.ell4_apply(idx4, w4, vals5), reps=3, warm=1)
    out["config5_" + name].update({
        "multilinear_regrid_ms_locate_weights_apply": ms_ml, "multilinear_Mpoints_per_s": m * m / ms_ml / 1e3,
        "apply_only_ms_per_frame": ms_ap,
        "apply_algorithmic_GBps": (64 + 8) * m * m / ms_ap / 1e6,  # idx4 + w4 read, one value written; gathers hit L2
        "nan_fraction": float(torch.isnan(res).double().mean())})
    del px, py, pxf, pyf, idx4, w4, res, vals5
# ---- fill(): 16 frames of 2048^2 with 8 % of the cells missing in blocks, 100 red-black iterations
from regridding_b200 import _fill
torch.cuda.empty_cache()
Tf, nf = 16, 2048
af = torch.rand((Tf, nf, nf), dtype=torch.float64, device=dev)
wf = torch.zeros((Tf, nf, nf), dtype=torch.bool, device=dev)
rngf = np.random.default_rng(3)
for t in range(Tf):
    for _ in range(80):
        j, i = rngf.integers(0, nf - 64, 2)
        wf[t, j:j + 64, i:i + 64] = True
n_missing = int(wf.sum())
ms_fill, _ = timed(lambda: _fill.gauss_seidel_2d_(af, wf, 100), reps=2, warm=1)
out["fill_gauss_seidel"] = {"frames": Tf, "shape": [nf, nf], "missing_cells": n_missing, "iterations": 100, "ms": ms_fill,
                            "cell_updates_per_s": n_missing * 100 / ms_fill * 1e3,
                            "note": "device-resident; includes building the six (colour, level) index lists"}
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    pathlib.Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
