"""Timeline (start, duration, gap to the previous end) of the launches of ONE band build, CUPTI via torch.profiler.
python tools/prof_band_timeline.py W r"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from regridding_b200 import _device, _parallel
from tests import cases
W, r = int(sys.argv[1]), int(sys.argv[2])
n = 2049
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (*gi, *co)]
lo, hi = _parallel.shard_range(n - 1, r, W)
def run():
    return _device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
for _ in range(5): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): run()
    torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
# last build only: from the last k_band_begin
starts = [i for i, e in enumerate(evs) if "k_band_begin" in e.name]
evs = evs[starts[-1]:]
t0 = evs[0].time_range.start
prev_end = t0
for e in evs:
    k = e.name.split("(")[0].replace("void ", "").replace("rg::", "")[:36]
    st, en = e.time_range.start - t0, e.time_range.end - t0
    print(f"{st:8.1f} +{en - st:7.1f}  gap {st - prev_end:6.1f}  {k}")
    prev_end = max(prev_end, en)
print(f"span {prev_end:.1f} us")
