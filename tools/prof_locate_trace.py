"""Per-kernel GPU durations of find_indices_2d at config 5 (CUPTI through torch.profiler). python tools/prof_locate_trace.py [scale]"""
import sys, pathlib, collections
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from regridding_b200 import _device
from tests import cases
dev = torch.device("cuda", 0)
gi5, _ = cases.benchmark_family(4096, distorted=True)
X, Y = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in gi5)
m = 8192
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
cx, cy = float(X.min() + X.max()) / 2, float(Y.min() + Y.max()) / 2
hx, hy = float(X.max() - X.min()) / 2 * scale, float(Y.max() - Y.min()) / 2 * scale
px = torch.linspace(cx - hx, cx + hx, m, dtype=torch.float64, device=dev)[:, None].expand(m, m).contiguous()
py = torch.linspace(cy - hy, cy + hy, m, dtype=torch.float64, device=dev)[None, :].expand(m, m).contiguous()
for _ in range(2): idx = _device.find_indices_2d(X, Y, px, py, -1)
torch.cuda.synchronize()
reps = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(reps): idx = _device.find_indices_2d(X, Y, px, py, -1)
    torch.cuda.synchronize()
order, tot = [], collections.defaultdict(float)
for e in sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start):
    k = e.name.split("(")[0].replace("void ", "").replace("rg::", "")[:40]
    if k not in tot: order.append(k)
    tot[k] += e.time_range.end - e.time_range.start
for k in order: print(f"{tot[k] / reps:9.1f} us  {k}")
print("fraction inside", float((idx >= 0).double().mean()))
