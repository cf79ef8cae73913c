"""One config-3 build (2049^2 vertices) + one apply, for ncu launch lists. Not a benchmark."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from regridding_b200 import _device
from tests import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2049
F = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
xi, yi, xo, yo = (torch.from_numpy(a).to(dev) for a in (*gi, *co))
vin = torch.rand((F, (n - 1) ** 2), dtype=torch.float64, device=dev)
for rep in range(reps):
    dw = _device.build_weights_2d(xi, yi, xo, yo, device=dev)
    csr = dw.csr()
    out = _device.apply_csr(csr, vin)
torch.cuda.synchronize()
print("done", dw.stats)
