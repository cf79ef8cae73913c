import sys; sys.path.insert(0, "/root/repo")
import torch
from regridding_b200 import _device
from tests import cases
n = 2049
dev = torch.device("cuda", 0)
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
dw = _device.build_weights_2d(*gi, *co, device=dev)
plan = dw.plan((n-1, n-1), (n-1, n-1))
ti = plan.tile_info.view(plan.n_tiles, -1)
cells = ti[:, 2].cpu(); nnz = ti[:, 3].cpu()
act = nnz > 0
print("tiles", plan.n_tiles, "active", int(act.sum()), "cells<=256 among active", float((cells[act] <= 256).float().mean()),
      "cells quantiles", torch.quantile(cells[act].float(), torch.tensor([0.1, 0.5, 0.9, 0.99, 1.0])).tolist())
