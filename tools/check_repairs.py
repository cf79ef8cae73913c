import sys; sys.path.insert(0, "/root/repo")
import torch
from regridding_b200 import _device
from tests import cases
dev = torch.device("cuda", 0)
for name in ("fam100", "dist129"):
    gi, go, _ = cases.case_2d(name)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    dw = _device.build_weights_2d(*[torch.from_numpy(a).to(dev) for a in (*gi, *co)], device=dev)
    print(name, dw.stats)
gi, go = cases.benchmark_family(2049, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
dw = _device.build_weights_2d(*[torch.from_numpy(a).to(dev) for a in (*gi, *co)], device=dev)
print("config3", dw.stats)
