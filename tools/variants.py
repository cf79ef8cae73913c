"""Time the planned apply for every library variant under regridding_b200/variants (development)."""
import os, subprocess, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
code = r'''
import sys, time; sys.path.insert(0, "%s")
import torch
from regridding_b200 import _device
from tests import cases
n, F = 2049, 256
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
dw = _device.build_weights_2d(*gi, *co, device=dev)
plan = dw.plan((n-1, n-1), (n-1, n-1))
vin = torch.rand((F, (n-1)**2), dtype=torch.float64, device=dev); out = torch.empty_like(vin)
for _ in range(3): _device.apply_planned(plan, vin, out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): _device.apply_planned(plan, vin, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = 8*F*2*(n-1)**2 + 12*dw.nnz + 4*((n-1)**2+1)
print("%%s: %%.3f ms  %%.0f GB/s  slots/nnz %%.3f" %% (sys.argv[1], ms, byt/ms/1e6, plan.slot_val.numel()/dw.nnz))
''' % ROOT
for lib in sorted((ROOT / "regridding_b200" / "variants").glob("lib_*.so")):
    env = dict(os.environ, REGRID_B200_LIB=str(lib))
    r = subprocess.run([sys.executable, "-c", code, lib.stem], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-500:])
