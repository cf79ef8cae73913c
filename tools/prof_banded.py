"""Phase timing of the banded multi-GPU build (torchrun). Development."""
import os, sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch, torch.distributed as dist
from regridding_b200 import _device, _parallel
from tests import cases
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 2049
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)
xi, yi, xo, yo = (torch.from_numpy(a).to(dev) for a in (*gi, *co))
def timeit(fn, reps=5):
    for _ in range(2): r = fn()
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    tt = torch.tensor([dt], device=dev, dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt) * 1e3, r
ms_full, _ = timeit(lambda: _device.build_weights_2d(xi, yi, xo, yo, device=dev))
ms_band, dw = timeit(lambda: _parallel.build_weights_2d_banded(xi, yi, xo, yo, replicate=False, device=dev))
ms_gather, _ = timeit(lambda: _parallel.allgather_concat([dw.indices_input, dw.indices_output, dw.values]))
ms_both, _ = timeit(lambda: _parallel.build_weights_2d_banded(xi, yi, xo, yo, replicate=True, device=dev))
if rank == 0:
    print(f"world {world}: full build {ms_full:.2f} ms | band only {ms_band:.2f} | gather only {ms_gather:.2f} | band + gather {ms_both:.2f}  (band nnz {dw.nnz})")
dist.destroy_process_group()
