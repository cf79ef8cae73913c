"""Band build (rg_build2d_band) at config 3 on ONE GPU: bit-equality of the concatenated bands with the full build and
the time of every band of a W-rank partition (what one rank of a W-GPU run does).  python tools/prof_band.py [W ...]"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from regridding_b200 import _device, _parallel
from tests import cases

n = 2049
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (*gi, *co)]

def timed(fn, reps=5, warm=2):
    for _ in range(warm): r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

ms_full, full = timed(lambda: _device.build_weights_2d(*t, device=dev))
print(f"full build {ms_full:.3f} ms  fragments {full.stats['fragments']} nnz {full.nnz}")
for W in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    parts, times = [], []
    for r in range(W):
        lo, hi = _parallel.shard_range(n - 1, r, W)
        def one():
            dw, status = _device.build2d_band_enqueue(*t, None, lo, hi, device=dev).finish()
            assert status in ("ok", "capacity"), status
            return dw, status
        one()
        ms, (dw, status) = timed(one)
        assert status == "ok"
        parts.append(dw); times.append(ms)
    same = (torch.equal(torch.cat([p.indices_input for p in parts]), full.indices_input) and
            torch.equal(torch.cat([p.indices_output for p in parts]), full.indices_output) and
            torch.equal(torch.cat([p.values for p in parts]), full.values))
    frs = [p.stats["fragments"] for p in parts]
    print(f"W={W}: bands equal full: {same}; per-band ms max {max(times):.3f} mean {np.mean(times):.3f} -> speedup {ms_full / max(times):.2f}x; "
          f"fragments {sum(frs)} (max band {max(frs)})")
