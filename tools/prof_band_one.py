"""One band of a W-rank band build at config 3, for ncu launch lists.  python tools/prof_band_one.py W r [reps]"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from regridding_b200 import _device, _parallel
from tests import cases
W, r = int(sys.argv[1]), int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = 2049
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (*gi, *co)]
lo, hi = _parallel.shard_range(n - 1, r, W)
for _ in range(reps):
    dw, status = _device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
print(status, dw.nnz if dw else None)
