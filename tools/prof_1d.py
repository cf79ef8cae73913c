"""Fused 1D conservative regrid at config-2 shape for ncu captures / timing. Not a benchmark."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from regridding_b200 import _device
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = 4097
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
base = torch.linspace(4000.0, 7000.0, n, dtype=torch.float64, device=dev)
xin = base * (1 + 1e-4 * torch.randn((S, 1), dtype=torch.float64, device=dev, generator=g)) + 0.3 * torch.sin(base / 500 + torch.rand((S, 1), dtype=torch.float64, device=dev, generator=g))
xout = torch.linspace(4001.0, 6999.0, n, dtype=torch.float64, device=dev) + 0.05 * torch.rand((S, 1), dtype=torch.float64, device=dev, generator=g)
vals = torch.rand((S, n - 1), dtype=torch.float64, device=dev, generator=g)
out = torch.empty((S, n - 1), dtype=torch.float64, device=dev)
for _ in range(3): _device.regrid1d_conservative(xin, xout, vals, out=out)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _device.regrid1d_conservative(xin, xout, vals, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
byt = S * 8 * (2 * n + 2 * (n - 1))
print(f"S={S}: min {min(ts):.3f} ms median {sorted(ts)[5]:.3f} ms  -> {byt/min(ts)/1e6:.0f} GB/s (min)")
