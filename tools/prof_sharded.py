"""Phase timing of the line-sharded 2D build.  Development.

1 process:   python tools/prof_sharded.py [W ...]   -- plays rank 0 of W on one GPU: times its walk share
             (count + fill) and the merge of band 0 (the fragments of the other ranks are produced untimed).
torchrun:    times the real thing (walk, all-to-all, merge) against the single-GPU build.
"""
import os, sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch, torch.distributed as dist
from regridding_b200 import _device, _parallel
from tests import cases

n = int(os.environ.get("N", "2049"))
gi, go = cases.benchmark_family(n, distorted=True)
co = cases.perturb_like_reference(go, (-1, -2), 42)


def ev_time(fn, reps=5, warm=2):
    r = None
    for _ in range(warm):
        r = None  # free the previous result first: the caching allocator then reuses its blocks (no cudaMalloc)
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = None
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


if "RANK" not in os.environ:
    dev = torch.device("cuda", 0)
    xi, yi, xo, yo = (torch.from_numpy(a).to(dev) for a in (*gi, *co))
    ms_full, full = ev_time(lambda: _device.build_weights_2d(xi, yi, xo, yo, device=dev))
    print(f"n={n}: single-GPU build {ms_full:.2f} ms, fragments {full.stats['fragments']}, nnz {full.nnz}")
    if os.environ.get("NCU"):  # launch list of rank 0 of 8: run under ncu --metrics gpu__time_duration.sum
        W = 8
        bounds = _parallel.band_bounds(n - 1, n - 1, W)
        parts = [_device.build2d_part_walk(xi, yi, xo, yo, None, r, W, bounds, device=dev) for r in range(W)]
        torch.cuda.synchronize()
        print("merge of band 0 from 8 sources", flush=True)
        cnt = _device.build2d_gather_counts([p.counts.data_ptr() for p in parts], bounds[1], dev)
        _device.build2d_merge(cnt, [p.frags.data_ptr() for p in parts], [p.band_offsets[1] for p in parts], 0,
                              parts[0].n_in, parts[0].n_out)
        torch.cuda.synchronize()
        sys.exit(0)
    for W in [int(a) for a in sys.argv[1:]] or [2, 4, 8]:
        bounds = _parallel.band_bounds(n - 1, n - 1, W)
        walk_ms, frag_share = [], []
        parts = []
        for r in range(W):
            ms, p = ev_time(lambda: _device.build2d_part_walk(xi, yi, xo, yo, None, r, W, bounds, device=dev), reps=3, warm=1)
            walk_ms.append(ms)
            frag_share.append(p.n_fragments)
            parts.append(p)
        merge_ms = []
        for d in range(W):
            lo, hi = bounds[d], bounds[d + 1]
            sizes = [p.band_offsets[d + 1] - p.band_offsets[d] for p in parts]
            ptrs = [p.frags.data_ptr() + 16 * p.band_offsets[d] for p in parts]

            def merge():
                cnt = _device.build2d_gather_counts([p.counts.data_ptr() + 4 * lo for p in parts], hi - lo, dev)
                return _device.build2d_merge(cnt, ptrs, sizes, lo, parts[0].n_in, parts[0].n_out)
            ms, dw = ev_time(merge, reps=3, warm=1)
            merge_ms.append(ms)
        print(f"W={W}: walk share per rank ms {['%.2f' % m for m in walk_ms]} (fragments {frag_share}); "
              f"merge per band ms {['%.2f' % m for m in merge_ms]}; "
              f"critical path without exchange {max(walk_ms) + max(merge_ms):.2f} ms "
              f"-> {ms_full / (max(walk_ms) + max(merge_ms)):.2f}x")
    sys.exit(0)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
xi, yi, xo, yo = (torch.from_numpy(a).to(dev) for a in (*gi, *co))


def timeit(fn, reps=5):
    for _ in range(3):
        r = fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt), r


ms_full, full = timeit(lambda: _device.build_weights_2d(xi, yi, xo, yo, device=dev))
for ex in (os.environ.get("EXCHANGE", "p2p,nccl")).split(","):
    try:
        ms_sh, dw = timeit(lambda: _parallel.build_weights_2d_sharded(xi, yi, xo, yo, replicate=False, device=dev, exchange=ex))
    except Exception as e:  # report and carry on with the other exchange
        if rank == 0:
            print(f"exchange {ex} failed: {type(e).__name__}: {e}")
        continue
    ms_rep, dwr = timeit(lambda: _parallel.build_weights_2d_sharded(xi, yi, xo, yo, replicate=True, device=dev, exchange=ex))
    ph = {}
    for _ in range(5):
        _parallel.build_weights_2d_sharded(xi, yi, xo, yo, replicate=False, device=dev, exchange=ex, phases=ph)
    ph = {k: round(v / 5, 3) for k, v in ph.items()}
    ok = bool(torch.equal(dwr.indices_input, full.indices_input) and torch.equal(dwr.indices_output, full.indices_output)
              and torch.equal(dwr.values, full.values))
    oks = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world {world} [{ex}]: full build {ms_full:.2f} ms | line-sharded {ms_sh:.2f} ms ({ms_full / ms_sh:.2f}x) | "
              f"+ all-gather {ms_rep:.2f} ms | replicated == single-GPU build on every rank: {bool(oks.item())}")
        print("   phases on rank 0 (ms):", ph)
dist.destroy_process_group()
