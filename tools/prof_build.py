"""Two config-3 builds for ncu launch lists. Not a benchmark."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from regridding_b200 import _device
from tests import cases
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2049
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
t = [torch.from_numpy(a).to(dev) for a in (*gi, *co)]
for _ in range(2):
    dw = _device.build_weights_2d(*t, device=dev)
torch.cuda.synchronize()
print("nnz", dw.nnz, dw.stats)
