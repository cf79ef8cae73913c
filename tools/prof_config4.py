"""Config 4 (per-slice grids) through _parallel.build_weights_2d_slices with a time breakdown. Development."""
import sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from regridding_b200 import _device, _parallel
from tests import cases
n, S = 2049, 8
dev = torch.device("cuda", 0)
grids = []
for f in range(S):
    gi, go = cases.benchmark_family(n, distorted=True, angle=0.4 + 0.002 * f, phase=float(f))
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    grids.append([torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (*gi, *co)])
host = [tuple(a.numpy() for a in g) for g in grids]
t = time.perf_counter(); x = [torch.from_numpy(a).pin_memory() for a in host[0]]; print("pin_memory() of pinned x4: %.2f ms" % ((time.perf_counter() - t) * 1e3), [a.is_pinned() for a in x])
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    built = _parallel.build_weights_2d_slices(host, device=dev)
    torch.cuda.synchronize(); print("slices (pinned host): %.1f ms for %d" % ((time.perf_counter() - t) * 1e3, S))
    del built
dv = [[a.to(dev) for a in g] for g in grids]
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    built = _device.build_weights_2d_batched(dv, device=dev)
    torch.cuda.synchronize(); print("batched (device resident): %.1f ms for %d" % ((time.perf_counter() - t) * 1e3, S))
    del built
pageable = [tuple(np.array(a) for a in h) for h in host]
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    built = _parallel.build_weights_2d_slices(pageable, device=dev)
    torch.cuda.synchronize(); print("slices (pageable host): %.1f ms for %d" % ((time.perf_counter() - t) * 1e3, S))
    del built
