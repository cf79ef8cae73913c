"""Config 5: find_indices of 8192^2 lattice points in a 4096^2-vertex grid; time + check against the oracle on a sample."""
import sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from regridding_b200 import _device
from tests import cases
from oracle import oracle
oracle.build(); oracle.set_num_threads()
dev = torch.device("cuda", 0)
gi5, _ = cases.benchmark_family(4096, distorted=True)
X, Y = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in gi5)
m = 8192
for name, scale in (("inside_0.7_bbox", 0.7), ("full_bbox", 1.0)):
    cx, cy = float(X.min() + X.max()) / 2, float(Y.min() + Y.max()) / 2
    hx, hy = float(X.max() - X.min()) / 2 * scale, float(Y.max() - Y.min()) / 2 * scale
    px = torch.linspace(cx - hx, cx + hx, m, dtype=torch.float64, device=dev)[:, None].expand(m, m).contiguous()
    py = torch.linspace(cy - hy, cy + hy, m, dtype=torch.float64, device=dev)[None, :].expand(m, m).contiguous()
    for _ in range(2): idx = _device.find_indices_2d(X, Y, px, py, -1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): idx = _device.find_indices_2d(X, Y, px, py, -1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    rng = np.random.default_rng(5)
    a, b = rng.integers(0, m, 50000), rng.integers(0, m, 50000)
    want = oracle.index_of_points(gi5[0], gi5[1], px[a, b].cpu().numpy(), py[a, b].cpu().numpy(), -1, "secant")
    ok = bool(np.array_equal(idx[a, b].cpu().numpy(), want))
    print(f"{name}: {ms:.3f} ms  {m*m/ms/1e3:.0f} Mpoints/s  {24*m*m/ms/1e6:.0f} GB/s = {24*m*m/ms/1e6/6537.3:.3f} of HBM  inside {float((idx>=0).double().mean()):.3f}  equals oracle on 50k: {ok}")
