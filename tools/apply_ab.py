"""A/B harness of the planned apply (development): every library variant under regridding_b200/variants/ plus the
product library, each in its own process: bit-equality against the generic CSR kernel on even / odd sizes, then the
time of a 256-frame apply at config 3.   python tools/apply_ab.py [--check-only]"""
import os, subprocess, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
code = r'''
import sys, time; sys.path.insert(0, "%s")
import numpy as np, torch
from regridding_b200 import _device
from tests import cases
dev = torch.device("cuda", 0)
name = sys.argv[1]; check = int(sys.argv[2])
def build(n, mx, my):
    gi, go = cases.benchmark_family(n, mx, my, distorted=True)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    return _device.build_weights_2d(*gi, *co, device=dev), (n-1, n-1), (mx-1, my-1)
if check:
    for (n, mx, my) in [(100, 120, 80), (258, 301, 222), (513, 513, 513), (300, 1100, 1100), (1100, 300, 301)]:
        dw, si, so = build(n, mx, my)
        plan = dw.plan(si, so)
        for F in (1, 8, 37, 64):
            x = torch.rand((F, dw.n_in), dtype=torch.float64, device=dev)
            a = _device.apply_csr(dw.csr(), x); b = _device.apply_planned(plan, x)
            ok = torch.equal(a, b)
            if not ok:
                bad = (a != b).nonzero()
                print(name, "MISMATCH", (n, mx, my), F, "generic tiles", plan.n_generic_tiles, "of", plan.n_tiles, "bad", bad.shape[0], bad[:5].tolist())
                break
        else:
            print(name, "ok", (n, mx, my), "generic tiles", plan.n_generic_tiles, "of", plan.n_tiles)
import os
n, F = 2049, int(os.environ.get('AB_FRAMES', '256'))
dw, si, so = build(n, n, n)
plan = dw.plan(si, so)
vin = torch.rand((F, (n-1)**2), dtype=torch.float64, device=dev); out = torch.empty_like(vin)
if check:
    a = _device.apply_csr(dw.csr(), vin[:24]); b = _device.apply_planned(plan, vin[:24])
    print(name, "config3 equal:", torch.equal(a, b), "generic tiles", plan.n_generic_tiles, "of", plan.n_tiles)
for _ in range(3): _device.apply_planned(plan, vin, out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): _device.apply_planned(plan, vin, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = 8*F*2*(n-1)**2 + 12*dw.nnz + 4*((n-1)**2+1)
print("%%s: %%.3f ms / %%d frames  %%.0f GB/s  frac %%.3f  slots/nnz %%.3f" %% (name, ms, F, byt/ms/1e6, byt/ms/1e6/6537.3, plan.slot_val.numel()/dw.nnz), flush=True)
''' % ROOT
libs = [ROOT / "regridding_b200" / "libregrid_b200.so"] + sorted((ROOT / "regridding_b200" / "variants").glob("lib_*.so"))
for lib in libs:
    env = dict(os.environ, REGRID_B200_LIB=str(lib))
    check = 0 if ("skip" in lib.stem) else 1
    try:
        r = subprocess.run([sys.executable, "-c", code, lib.stem, str(check)], env=env, capture_output=True, text=True,
                           timeout=100)
    except subprocess.TimeoutExpired as e:
        print(lib.stem, "TIMEOUT", (e.stdout or b"")[-800:], (e.stderr or b"")[-800:])
        continue
    print(r.stdout.strip() or "(no output)")
    if r.returncode != 0:
        print(lib.stem, "FAILED rc", r.returncode, r.stderr[-1500:])
