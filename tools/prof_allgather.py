"""Times the pieces of the replicated-build all-gather (torchrun). Development."""
import os, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
from regridding_b200 import _parallel
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 15_025_254 // world + rank * 1000
ts = [torch.arange(n, dtype=torch.int64, device=dev), torch.arange(n, dtype=torch.int64, device=dev) * 2,
      torch.rand(n, dtype=torch.float64, device=dev)]

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def steps():
    m = [("start", ev())]
    nn = torch.tensor([n], dtype=torch.int64, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, nn)
    ch = counts.cpu().tolist(); m.append(("counts+sync", ev()))
    most = max(ch)
    packed = torch.empty((3, most), dtype=torch.int64, device=dev)
    for q, t in enumerate(ts):
        packed[q, :n] = t.view(torch.int64)
    m.append(("pack", ev()))
    gathered = torch.empty((world, 3, most), dtype=torch.int64, device=dev); m.append(("alloc", ev()))
    dist.all_gather_into_tensor(gathered, packed); m.append(("all_gather", ev()))
    outs = [torch.cat([gathered[r, q, :ch[r]] for r in range(world)]) for q in range(3)]; m.append(("compact", ev()))
    torch.cuda.synchronize()
    return {b[0]: round(a[1].elapsed_time(b[1]), 3) for a, b in zip(m[:-1], m[1:])}

for _ in range(3):
    r = steps()
dist.barrier(); torch.cuda.synchronize()
r = steps()
if rank == 0:
    print("steps (ms):", r)
import time
for _ in range(2):
    _parallel.allgather_concat(ts)
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    _parallel.allgather_concat(ts)
torch.cuda.synchronize()
if rank == 0:
    print("allgather_concat wall ms", (time.perf_counter() - t0) / 5 * 1e3)
dist.destroy_process_group()
