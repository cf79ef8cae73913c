"""Raw pinned-memory PCIe rates of this box (H2D, D2H, both at once): the ceiling of the e2e figure. Development."""
import time, torch
dev = torch.device("cuda", 0)
nb = 1 << 30
h_in = torch.empty(nb, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(nb, dtype=torch.uint8, device=dev)
d_out = torch.empty(nb, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def run(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return reps * nb * (h2d + d2h) / (time.perf_counter() - t) / 1e9
run(1, 1)
print(f"H2D {run(1,0):.1f} GB/s  D2H {run(0,1):.1f} GB/s  both {run(1,1):.1f} GB/s (sum)")
import os
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print(open("/proc/meminfo").read().split("\n")[0])
    import subprocess; print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
except Exception as e: print(e)
