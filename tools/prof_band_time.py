"""Wall time per band of a W-rank band build at config 3 on ONE GPU (no collective): CUDA events around `reps`
builds, each with its finish() (the one host sync), like bench.py times them.  python tools/prof_band_time.py W [r ...]"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from regridding_b200 import _device, _parallel
from tests import cases
W = int(sys.argv[1]); ranks = [int(a) for a in sys.argv[2:]] or list(range(W))
n = 2049
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (*gi, *co)]
for _ in range(3): full = _device.build_weights_2d(*t, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): full = _device.build_weights_2d(*t, device=dev)
e1.record(); torch.cuda.synchronize()
single = e0.elapsed_time(e1) / 10
print(f"single-GPU build {single:.3f} ms")
worst = 0.0
for r in ranks:
    lo, hi = _parallel.shard_range(n - 1, r, W)
    for _ in range(3): dw, status = _device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
    e0.record()
    for _ in range(10): dw, status = _device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    worst = max(worst, ms)
    sel = (full.indices_input >= lo * (n - 1)) & (full.indices_input < hi * (n - 1))
    same = torch.equal(dw.indices_input, full.indices_input[sel]) and torch.equal(dw.indices_output, full.indices_output[sel]) and torch.equal(dw.values, full.values[sel])
    print(f"band {r}/{W}: {ms:.3f} ms  status {status}  nnz {dw.nnz}  equal {same}")
print(f"slowest band {worst:.3f} ms -> {single / worst:.2f}x of the single-GPU build (without the status all-reduce)")
