"""End-to-end regrid_from_weights (pinned host buffers) against the pipeline chunk size. Development."""
import sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import regridding_b200 as rg
from regridding_b200 import _device, _regrid, _cache
from tests import cases
n, Fe = 2049, 125
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
dw = _device.build_weights_2d(*[torch.from_numpy(a).to(dev) for a in (*gi, *co)], device=dev)
w = np.empty((), dtype=object); w[()] = dw.to_host(); _cache.remember(w[()], dw)
pin_in = torch.empty((Fe, n - 1, n - 1), dtype=torch.float64, pin_memory=True); pin_in.uniform_(0, 1)
pin_out = torch.empty((Fe, n - 1, n - 1), dtype=torch.float64, pin_memory=True)
hi, ho = pin_in.numpy(), pin_out.numpy()
byt = 8 * Fe * 2 * (n - 1) ** 2
for chunk_mb, ring in [(256, 3), (128, 3), (64, 3), (128, 4), (64, 6), (32, 6)]:
    _regrid._CHUNK_BYTES = chunk_mb << 20; _regrid._RING = ring
    for _ in range(2): rg.regrid_from_weights(w, (n - 1, n - 1), (n - 1, n - 1), hi, values_output=ho)
    t = time.perf_counter()
    for _ in range(3): rg.regrid_from_weights(w, (n - 1, n - 1), (n - 1, n - 1), hi, values_output=ho)
    dt = (time.perf_counter() - t) / 3
    print(f"chunk {chunk_mb} MB ring {ring}: {dt*1e3:.1f} ms  {byt/dt/1e9:.1f} GB/s")
