"""Planned apply at config-3 size for ncu captures. Not a benchmark."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from regridding_b200 import _device
from tests import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2049
F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
dw = _device.build_weights_2d(*gi, *co, device=dev)
plan = dw.plan((n - 1, n - 1), (n - 1, n - 1))
vin = torch.rand((F, (n - 1) ** 2), dtype=torch.float64, device=dev)
out = torch.empty_like(vin)
for rep in range(reps):
    _device.apply_planned(plan, vin, out)
torch.cuda.synchronize()
print("done", plan.n_generic_tiles, plan.n_tiles)
