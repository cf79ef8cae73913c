"""Config-1 sized build timing (launch-latency bound). Development."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from regridding_b200 import _device
from tests import cases
dev = torch.device("cuda", 0)
gi, go, _ = cases.case_2d("fam100")
co = cases.perturb_like_reference(go, (-1, -2), 42)
t = [torch.from_numpy(a).to(dev) for a in (*gi, *co)]
for _ in range(5): _device.build_weights_2d(*t, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): _device.build_weights_2d(*t, device=dev)
e1.record(); torch.cuda.synchronize()
print("config-1 build ms", e0.elapsed_time(e1) / 20)
