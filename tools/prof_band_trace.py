"""Per-kernel GPU durations of one band build in stream order (CUPTI activity records through torch.profiler: not
serialised, warm caches -- unlike an ncu launch list).  python tools/prof_band_trace.py W r [reps]   (W = 1: the
single-GPU build)"""
import sys, pathlib, collections
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from regridding_b200 import _device, _parallel
from tests import cases
W, r = int(sys.argv[1]), int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n = 2049
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (*gi, *co)]
lo, hi = _parallel.shard_range(n - 1, r, W)
def run():
    if W == 1:
        return _device.build_weights_2d(*t, device=dev)
    return _device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
for _ in range(3): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(reps): run()
    torch.cuda.synchronize()
order, tot, cnt = [], collections.defaultdict(float), collections.defaultdict(int)
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
for e in evs:
    k = e.name.split("(")[0].replace("void ", "").replace("rg::", "")[:40]
    if k not in tot: order.append(k)
    tot[k] += e.time_range.end - e.time_range.start; cnt[k] += 1
span = (evs[-1].time_range.end - evs[0].time_range.start) / reps
s = 0.0
for k in order:
    print(f"{tot[k] / reps:9.1f} us  x{cnt[k] / reps:4.1f}  {k}")
    s += tot[k] / reps
print(f"sum of kernels {s:.1f} us, span per build {span:.1f} us")
