#!/usr/bin/env python
"""
bench.py -- BASELINE.json metric on BASELINE.json config 3 (the largest single-GPU
configuration the metric is quoted on):

    2D conservative weights build for a 2048x2048-cell distorted curvilinear grid onto a
    2048x2048-cell rectilinear grid, then regrid_from_weights over 1000 frames sharing the
    weights.

One "step" = one pass of regrid_from_weights over the F frames resident on this GPU
(primary metric, regrid_from_weights GB/s, HBM roofline); the weights build (Mcells/s) is
measured in the same run and reported in the "build" object of the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Multi-GPU: frames are independent units, so every rank applies the shared weights to its
own F frames with NO data-path collective ("scaling": "weak"); the 2D build is additionally
timed as ONE build sharded over the ranks (strong scaling, "build.sharded": rank r builds the band of input
rows it owns and walks only the sweep segments that can reach it -- no fragments are exchanged, no collective; "replicated_ms" adds the all-gather that leaves the full matrix on every rank).

--impl reference: the CPU oracle (a C port of the reference's algorithm, OpenMP over the
reference's own prange loops) on this box's host cores, same metric / config, bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "regrid_from_weights_GBps"
UNIT = "GB/s"
WORKLOAD = ("config3: distorted curvilinear 2049x2049 vertices (2048^2 cells) -> rectilinear 2049x2049, "
            "weights build + regrid_from_weights over F frames sharing the weights")


def grids(n: int):
    from tests import cases

    gi, go = cases.benchmark_family(n, distorted=True)
    return gi, go


def apply_bytes(F: int, n_in: int, n_out: int, nnz: int) -> int:
    """ALGORITHMIC bytes of one apply: values read once + written once, weights once."""
    return 8 * F * (n_in + n_out) + 12 * nnz + 4 * (n_out + 1)


def build_bytes(v_in: int, v_out: int, nnz: int) -> int:
    return 16 * (v_in + v_out) + 24 * nnz


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the committed
# `ncu --set full` capture of THIS kernel, scaled to the launch the bench times; None until measured.
# profiles/r2_apply_ncu_summary.txt (k_apply_bulk, 256-frame launch at config 3): 9.579 GB read + 8.540 GB written
# = 70.78 MB per frame (1.04x the algorithmic bytes; footprint halos that miss L2 + the slot arrays).
TRAFFIC_NCU: dict = {"apply_bytes_per_frame": (9.578994e9 + 8.539667e9) / 256,
                     "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one 256-frame launch of "
                               "rg::k_apply_bulk, scaled to this launch's frames (profiles/r2_apply_ncu_summary.txt)"}
# fp64-pipe utilisation of the build's walk kernels from the committed ncu capture (north_star: "fp64-pipe utilisation
# for the build"): sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed, profiles/r2_build_ncu_summary.txt
BUILD_NCU: dict = {"fp64_pipe_pct": {"k_walk_count": 18.78, "k_walk_emit": 4.90, "k_vertex_guess": 43.79},
                   "issue_slots_pct": {"k_walk_count": 70.7, "k_walk_emit": 35.0, "k_bucket_sort": 82.6, "k_bucket_emit": 24.6},
                   "dram_bytes": {"k_walk_count": 1.989e9, "k_walk_emit": 2.817e9, "k_bucket_sort": 1.919e9, "k_bucket_emit": 1.477e9},
                   "source": "profiles/r2_build_ncu_summary.txt (ncu --set full of one config-3 build)"}
FLOPS_PER_PIECE = 64  # 3 edge tests x 17 + 6 intersection point + 4 area + 2 weight divides + 1 negate


# The reference's own Numba implementation (as shipped, warm JIT cache) measured in the build container
# (8 vCPU, numba 0.65, OpenMP layer; SURVEY.md section 6 / BASELINE.md): printed beside the port's numbers because the
# CPU arm of this bench is the C/OpenMP port of the same algorithm, which is several times faster than Numba.
NUMBA_FIGURES = {"apply_GBps": 6.0, "build_Mcells_per_s": 0.058, "cores": 8,
                 "source": "SURVEY.md section 6 (2048^2 cells: sweep 64.7 s + coalesce 7.2 s; apply 16 frames 0.240 s)"}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None
        self.active = False

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):   # development: rule the sampler out as a disturbance
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            if self.active:
                self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU legs (oracle = port of the reference's algorithm; test/bench infrastructure)
# ---------------------------------------------------------------------------


def cpu_apply_sample(ii, io, v, n_in, n_out, frames, reps=3):
    from oracle import oracle

    vals = np.random.default_rng(0).random((frames, n_in))
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        oracle.regrid_from_weights(ii, io, v, vals, n_out)
        best = min(best, time.perf_counter() - t)
    return apply_bytes(frames, n_in, n_out, v.size) / best / 1e9, best


def cpu_build_sample(n):
    from oracle import oracle
    from tests import cases

    gi, go = grids(n)
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    t = time.perf_counter()
    raw = oracle.weights_conservative_2d(gi, co)
    t_sweep = time.perf_counter() - t
    t = time.perf_counter()
    tri = oracle.coalesce(*raw)
    t_coal = time.perf_counter() - t
    return tri, t_sweep, t_coal


def run_reference(args):
    """Reference arm: the oracle port on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    from tests import cases

    oracle.build()
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: use every core this process may run on
    cores = oracle.set_num_threads()
    n = args.n
    n_in = n_out = (n - 1) ** 2
    tri, t_sweep, t_coal = cpu_build_sample(n)
    weights_sha = cases.sha(*tri)
    frames = args.ref_frames
    vals = np.random.default_rng(0).random((frames, n_in))
    for _ in range(args.warmup):
        oracle.regrid_from_weights(*tri, vals, n_out)
    t = time.perf_counter()
    for _ in range(args.steps):
        oracle.regrid_from_weights(*tri, vals, n_out)
    dt = (time.perf_counter() - t) / args.steps
    value = apply_bytes(frames, n_in, n_out, tri[2].size) / dt / 1e9
    sample = f"{frames} of the F frames per step, full 2048^2 weights (nnz {tri[2].size})"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": frames, "grid_vertices": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "build": {"metric": "conservative_2d_weights_build_Mcells_per_s",
                  "value": n_in / (t_sweep + t_coal) / 1e6, "unit": "Mcells/s",
                  "sweep_s": t_sweep, "coalesce_s": t_coal, "cores": cores, "kind": "port",
                  "sample": "full 2048^2-cell build, 1 repetition"},
        "gpu_launches": 0,
        # SHA-256 of the reference algorithm's (ii, io, v) for this config: the GPU arm prints the same key
        "weights_sha256": weights_sha,
        "reference_numba_on_8_vcpu": NUMBA_FIGURES,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------


def run_ours(args):
    import torch
    import torch.distributed as dist

    import regridding_b200 as rg
    from regridding_b200 import _device, _parallel, _util
    from tests import cases

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version / debug lines must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    F = args.frames
    K, W = args.steps, max(args.warmup, 3)
    n_in = n_out = (n - 1) ** 2
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- inputs (synthetic, host-generated, seeds fixed) -------------------------------
    gi, go = grids(n)
    t0 = time.perf_counter()
    co = cases.perturb_like_reference(go, (-1, -2), _util.SEED_DEFAULT)  # the reference's host jitter
    t_perturb = time.perf_counter() - t0
    xi, yi, xo, yo = (torch.from_numpy(a).to(dev) for a in (*gi, *co))

    # ---- build: device-resident coordinates -> device-resident public triplets -----------
    exchange = {"mode": "band"}

    def build_once(sharded: bool, replicate: bool = False):
        if sharded and world > 1:
            return _parallel.build_weights_2d_sharded(xi, yi, xo, yo, replicate=replicate, device=dev,
                                                      exchange=exchange["mode"])
        return _device.build_weights_2d(xi, yi, xo, yo, device=dev)

    def time_build(sharded: bool, replicate: bool = False):
        for _ in range(W):
            dw = build_once(sharded, replicate)
        barrier()
        sampler.active = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            dw = build_once(sharded, replicate)
        e1.record()
        torch.cuda.synchronize(dev)
        sampler.active = False
        ms = max_over_ranks(e0.elapsed_time(e1) / K)
        barrier()
        return dw, ms

    # The FIRST build of a shape takes the standard pipeline (two walks: the bucket offsets must be exact before anything
    # is emitted); a REPEAT build knows the longest bucket of the first and walks every segment once
    # (rg_build2d_band_onewalk, the whole grid as one band, replayed as a CUDA graph).  `build_weights_2d` picks the
    # path itself; the first-build figure is measured with the one-walk path switched off.
    os.environ["REGRID_B200_BAND_TWO_WALKS"] = "1"
    try:
        dw_first, first_build_ms = time_build(False)
    finally:
        del os.environ["REGRID_B200_BAND_TWO_WALKS"]
    dw, build_ms = time_build(False)
    assert torch.equal(dw_first.indices_input, dw.indices_input) and torch.equal(dw_first.indices_output, dw.indices_output) and \
        torch.equal(dw_first.values, dw.values), "one-walk repeat build differs from the standard build"
    build_stats = dict(dw_first.stats or {})
    del dw_first
    one_walk_ms = build_ms
    sharded_ms = replicated_ms = sharded_first_ms = None
    sharded_equal = None
    if world > 1:
        # agree on a fallback together should the default exchange be unavailable on some rank
        ok = 1
        try:
            build_once(True)
        except Exception as e:  # noqa: BLE001
            ok = 0
            sys.stderr.write(f"[rank {rank}] {exchange['mode']} build unavailable ({type(e).__name__}: {e}); using NCCL all-to-all\n")
        t = torch.tensor([ok], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 0:
            exchange["mode"] = "nccl"
        dw_band, sharded_ms = time_build(True)
        # the FIRST build of a shape cannot know the bucket capacity: two walks (forced here for every timed build)
        os.environ["REGRID_B200_BAND_TWO_WALKS"] = "1"
        try:
            _, sharded_first_ms = time_build(True)
        finally:
            del os.environ["REGRID_B200_BAND_TWO_WALKS"]
        dw_rep, replicated_ms = time_build(True, replicate=True)
        # parity before any number is reported: this rank's band and the replicated matrix equal the single-GPU
        # build bit for bit (indices AND weights)
        lo, hi = _parallel.band_cells(n - 1, n - 1, rank, world)
        sel = (dw.indices_input >= lo) & (dw.indices_input < hi)
        same = int(torch.equal(dw_band.indices_input, dw.indices_input[sel]) and
                   torch.equal(dw_band.indices_output, dw.indices_output[sel]) and
                   torch.equal(dw_band.values, dw.values[sel]) and
                   torch.equal(dw_rep.indices_input, dw.indices_input) and
                   torch.equal(dw_rep.indices_output, dw.indices_output) and
                   torch.equal(dw_rep.values, dw.values))
        t = torch.tensor([same], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        sharded_equal = bool(int(t.item()))
        assert sharded_equal, "sharded / replicated build differs from the single-GPU build"
        del dw_band, sel
        dw = dw_rep
    nnz = dw.nnz
    csr = dw.csr()
    plan = dw.plan((n - 1, n - 1), (n - 1, n - 1))  # per-tile footprints of the shared-memory staged apply
    torch.cuda.synchronize(dev)

    # ---- apply: F frames resident in HBM ---------------------------------------------
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    vin = torch.empty((F, n_in), dtype=torch.float64, device=dev)
    chunk = 50
    for f in range(0, F, chunk):
        vin[f:f + chunk].uniform_(0.0, 1.0, generator=gen)
    vout = torch.empty((F, n_out), dtype=torch.float64, device=dev)
    for _ in range(W):
        _device.apply_planned(plan, vin, vout)
    barrier()
    sampler.active = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        _device.apply_planned(plan, vin, vout)
    e1.record()
    torch.cuda.synchronize(dev)
    sampler.active = False
    apply_ms = max_over_ranks(e0.elapsed_time(e1) / K)
    barrier()
    abytes = apply_bytes(F, n_in, n_out, nnz)
    value = world * abytes / (apply_ms * 1e-3) / 1e9
    per_gpu = abytes / (apply_ms * 1e-3) / 1e9
    checksum = float(vout[:2].sum().item())
    del vin, vout
    torch.cuda.empty_cache()

    # ---- e2e: reference-facing public API with HOST buffers ------------------------------
    Fe = min(args.e2e_frames, F)
    weights_host = np.empty((), dtype=object)
    weights_host[()] = dw.to_host()
    from regridding_b200 import _cache

    _cache.remember(weights_host[()], dw)
    weights_sha = cases.sha(*weights_host[()]) if rank == 0 else None
    shape_in = shape_out = (n - 1, n - 1)
    pin_in = torch.empty((Fe, n - 1, n - 1), dtype=torch.float64, pin_memory=True)
    pin_in.uniform_(0.0, 1.0)
    pin_out = torch.empty((Fe, n - 1, n - 1), dtype=torch.float64, pin_memory=True)
    host_in, host_out = pin_in.numpy(), pin_out.numpy()
    for _ in range(2):
        rg.regrid_from_weights(weights_host, shape_in, shape_out, host_in, values_output=host_out)
    barrier()
    t0 = time.perf_counter()
    Ke = max(2, min(K, 5))
    for _ in range(Ke):
        rg.regrid_from_weights(weights_host, shape_in, shape_out, host_in, values_output=host_out)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks((time.perf_counter() - t0) / Ke)
    e2e_value = world * apply_bytes(Fe, n_in, n_out, nnz) / e2e_s / 1e9
    del pin_in, pin_out

    # e2e build through the public API: host coordinates -> host triplets (perturb + H2D + build + D2H)
    t0 = time.perf_counter()
    Wh = rg.weights(gi, go, method="conservative")
    torch.cuda.synchronize(dev)
    e2e_build_s = max_over_ranks(time.perf_counter() - t0)
    assert Wh[0][()][2].size == nnz

    if rank == 0:
        sampler.stop()
    peak, peak_src = hbm_peak()
    clocks = sampler.summary() if rank == 0 else None

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) ----------------------
    cpu = None
    cpu_build = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle

        oracle.build()
        oracle.set_num_threads()
        ii_h, io_h, v_h = weights_host[()]
        frames_cpu = args.cpu_frames
        cpu_val, cpu_t = cpu_apply_sample(ii_h, io_h, v_h, n_in, n_out, frames_cpu)
        cpu = {"value": cpu_val, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
               "sample": f"{frames_cpu} frames with the full 2048^2 weights, best of 3 ({cpu_t:.2f} s each)"}
        nb = args.cpu_build_n
        _, ts, tc = cpu_build_sample(nb)
        cpu_build = {"value": (nb - 1) ** 2 / (ts + tc) / 1e6, "unit": "Mcells/s", "cores": oracle.num_threads(),
                     "kind": "port",
                     "sample": f"{nb - 1}^2-cell grid of the same family (sweep {ts:.1f} s + coalesce {tc:.1f} s); "
                               "the CPU algorithm is super-linear in the grid size, so this flatters it"}

    extras = None
    if not args.no_extras:
        torch.cuda.empty_cache()
        extras = extra_configs(dev, rank, world, peak, barrier, max_over_ranks, with_cpu=(rank == 0 and world == 1 and not args.no_cpu))

    # fp64 FMA-chain throughput of this device: the denominator of the build's fp64 roofline
    import ctypes

    from regridding_b200 import _lib

    tf = ctypes.c_double()
    _lib.check(_lib.load().rg_measure_fp64_peak(dev.index, torch.cuda.current_stream(dev).cuda_stream, ctypes.byref(tf)),
               "rg_measure_fp64_peak")
    fp64_peak = float(tf.value)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    v_in, v_out = n * n, n * n
    bbytes = build_bytes(v_in, v_out, nnz)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": apply_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": F, "grid_vertices": n, "nnz": nnz,
                   "l2": "inputs (33.5 GB in + 33.5 GB out per step at F=1000) are far larger than L2; no flush needed",
                   "parallelism": f"frames x{world} (no collective)"},
        "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak,
                     "traffic": (TRAFFIC_NCU["apply_bytes_per_frame"] * F if n == 2049 else None),
                     "traffic_source": TRAFFIC_NCU["source"],
                     "peak_source": peak_src, "kernel": "rg::k_apply_bulk",
                     "algorithmic_bytes_per_launch": abytes},
        "e2e": {"value": e2e_value, "unit": UNIT, "frames": Fe,
                "h2d_bytes_per_step": 8 * Fe * n_in, "d2h_bytes_per_step": 8 * Fe * n_out,
                "api": "regridding_b200.regrid_from_weights(numpy in pinned host memory -> numpy)"},
        "gpu_launches": K * _device.LAUNCHES_APPLY * ((F + 512 * 65535 - 1) // (512 * 65535)),
        "clocks": clocks,
        "cpu_baseline": cpu,
        "build": {
            "metric": "conservative_2d_weights_build_Mcells_per_s", "unit": "Mcells/s",
            "value": n_in / (build_ms * 1e-3) / 1e6, "ms": build_ms,
            "fragments": build_stats.get("fragments"),
            "repaired_segments": build_stats.get("repaired_segments"),
            "scope": "perturbed device-resident coordinates -> device-resident sorted-unique (ii, io, v)",
            "gpu_launches_per_build": _device.LAUNCHES_BUILD2D,
            "roofline": {"bound": "hbm", "achieved": bbytes / (build_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bbytes / (build_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": bbytes,
                         "note": "the build is fp64-latency / divergence bound, not HBM bound; see DESIGN.md"},
            "roofline_fp64": {
                "bound": "fp64", "unit": "TFLOP/s", "peak": fp64_peak,
                "peak_source": "rg_measure_fp64_peak: 8 independent DFMA chains per thread on all SMs, measured in this run",
                "flops": FLOPS_PER_PIECE * (build_stats.get("fragments") or 0) // 2,
                "flops_model": "64 fp64 flops per piece (SURVEY.md 8d), pieces = raw fragments / 2",
                "achieved": FLOPS_PER_PIECE * ((build_stats.get("fragments") or 0) // 2) / (build_ms * 1e-3) / 1e12,
                "frac": FLOPS_PER_PIECE * ((build_stats.get("fragments") or 0) // 2) / (build_ms * 1e-3) / 1e12 / fp64_peak,
                "ncu_fp64_pipe_pct": BUILD_NCU["fp64_pipe_pct"], "ncu_issue_slots_pct": BUILD_NCU["issue_slots_pct"],
                "ncu_dram_bytes": BUILD_NCU["dram_bytes"], "ncu_source": BUILD_NCU["source"]},
            # config 4 (every orthogonal slice carries its own grid): slices shard across ranks with no collective,
            # every rank runs full builds of its own slices -> aggregate = ranks x the per-rank rate (max over ranks)
            "per_slice_sharded": {"n_gpus": world, "value": world * n_in / (first_build_ms * 1e-3) / 1e6, "unit": "Mcells/s",
                                  "scaling": "weak", "collective": None},
            "first_build": {
                "ms": first_build_ms, "value": n_in / (first_build_ms * 1e-3) / 1e6, "unit": "Mcells/s",
                "equals_repeat_build_bitwise": True,
                "note": "`build.ms` / `build.value` are REPEAT builds of the shape (build_weights_2d routes them through "
                        "rg_build2d_band_onewalk: the longest bucket of the first build sizes fixed-capacity buckets, every "
                        "segment is walked once); this is the standard pipeline a first build takes (two walks), and the "
                        "roofline / ncu figures of this object were taken on it"},
            "sharded": None if sharded_ms is None else {
                "n_gpus": world, "ms": sharded_ms, "value": n_in / (sharded_ms * 1e-3) / 1e6, "unit": "Mcells/s",
                "speedup_vs_this_runs_1gpu_build": build_ms / sharded_ms,
                "first_build_ms": sharded_first_ms,
                "like_for_like": {
                    "repeat_builds": {"one_gpu_ms": one_walk_ms, "n_gpu_ms": sharded_ms, "speedup": one_walk_ms / sharded_ms},
                    "first_builds": {"one_gpu_ms": first_build_ms, "n_gpu_ms": sharded_first_ms,
                                     "speedup": first_build_ms / sharded_first_ms},
                    "mixed": {"one_gpu_first_build_ms": first_build_ms, "n_gpu_repeat_ms": sharded_ms,
                              "speedup": first_build_ms / sharded_ms},
                    "note": "`ms` / `speedup_vs_this_runs_1gpu_build` compare repeat builds with repeat builds (one walk on "
                            "both sides); `mixed` divides the standard single-GPU build by the N-GPU repeat build"},
                "scaling": "strong", "exchange": exchange["mode"],
                "equals_single_gpu_build_bitwise_on_every_rank": sharded_equal,
                "result": "every rank holds its input-row band of the public triplets; exchange=band: every rank walks only "
                          "the sweep segments that can reach its band and verifies its own walk states: no fragments are "
                          "exchanged, no collective",
                "replicated_ms": replicated_ms,
                "replicated_collective": "all-gather of the band triplets (full matrix on every rank)"},
            "e2e": {"value": n_in / e2e_build_s / 1e6, "unit": "Mcells/s", "seconds": e2e_build_s,
                    "host_perturb_seconds": t_perturb,
                    "h2d_bytes": 16 * (v_in + v_out), "d2h_bytes": 24 * nnz,
                    "api": "regridding_b200.weights(method='conservative') host coords -> host triplets"},
            "cpu_baseline": cpu_build,
        },
        "configs": extras,
        "checksum": checksum,
        # SHA-256 of the public (ii, io, v): `--impl reference` prints the same key for the CPU oracle's triplets
        "weights_sha256": weights_sha,
        "reference_numba_on_8_vcpu": NUMBA_FIGURES,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ---------------------------------------------------------------------------
# the other BASELINE.json configurations (1, 2, 4, 5): extra keys of the same JSON line
# ---------------------------------------------------------------------------


def extra_configs(dev, rank, world, peak, barrier, max_over_ranks, with_cpu):
    """One measured object per BASELINE config besides config 3; every rank does the same amount of work (weak
    scaling, aggregate = world x per-rank rate by the slowest rank) except config 4, whose 8 x world frames are
    sharded over the ranks (each frame carries its own grid: no collective)."""
    import torch

    import regridding_b200 as rg
    from regridding_b200 import _device
    from tests import cases

    out = {}

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            r = fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return max_over_ranks(e0.elapsed_time(e1) / reps), r

    def roof(bytes_, ms):
        a = bytes_ / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "algorithmic_bytes": bytes_}

    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731

    # ---- config 1: 100x100 vertices -> 120x80, one image -------------------------------------------------
    gi, go, _ = cases.case_2d("fam100")
    co = cases.perturb_like_reference(go, (-1, -2), 42)
    t = [T(a) for a in (*gi, *co)]
    ms_b, dw = timed(lambda: _device.build_weights_2d(*t, device=dev))
    plan = dw.plan((99, 99), (119, 79))
    x = torch.rand((1, dw.n_in), dtype=torch.float64, device=dev)
    ms_a, _ = timed(lambda: _device.apply_planned(plan, x), reps=20)
    vals1 = np.random.default_rng(0).random((99, 99))
    rg.regrid(gi, go, vals1, method="conservative")
    t0 = time.perf_counter()
    rg.regrid(gi, go, vals1, method="conservative")
    t_api = time.perf_counter() - t0
    out["config1"] = {"workload": "100x100 vertices -> 120x80 vertices, one image", "build_ms": ms_b, "apply_ms_1_frame": ms_a,
                      "nnz": dw.nnz, "staged_tiles": plan.n_tiles - plan.n_generic_tiles, "tiles": plan.n_tiles,
                      "regrid_api_s_host_to_host": t_api, "build_Mcells_per_s": dw.n_in / ms_b / 1e3,
                      "reference_numba_regrid_s": 0.0317}

    # ---- config 2: S spectra x 4096 bins with per-spectrum grids, fused 1D conservative regrid -----------
    S, n = 262144, 4097
    g = torch.Generator(device=dev)
    g.manual_seed(rank)
    base = torch.linspace(4000.0, 7000.0, n, dtype=torch.float64, device=dev)
    xin = base * (1 + 1e-4 * torch.randn((S, 1), dtype=torch.float64, device=dev, generator=g)) + \
        0.3 * torch.sin(base / 500 + torch.rand((S, 1), dtype=torch.float64, device=dev, generator=g))
    xout = torch.linspace(4001.0, 6999.0, n, dtype=torch.float64, device=dev) + \
        0.05 * torch.rand((S, 1), dtype=torch.float64, device=dev, generator=g)
    vals = torch.rand((S, n - 1), dtype=torch.float64, device=dev, generator=g)
    res = torch.empty((S, n - 1), dtype=torch.float64, device=dev)
    ms, _ = timed(lambda: _device.regrid1d_conservative(xin, xout, vals, out=res), reps=5)
    byt = S * (8 * n + 8 * n + 8 * (n - 1) + 8 * (n - 1))
    c2 = {"workload": f"{S} spectra per GPU x 4096 bins, per-spectrum wavelength grids, fused conservative regrid "
                      "(inputs resident in HBM)",
          "ms": ms, "spectra_per_s": world * S / ms * 1e3, "roofline": roof(byt, ms), "n_gpus": world,
          "full_1M_spectra_s_extrapolated_per_gpu": ms * (1e6 / S) / 1e3,
          "reference_numba_spectra_per_s_8_vcpu": 592}
    if with_cpu:
        from oracle import oracle

        Sc = 2048
        xi_h, xo_h, v_h = xin[:Sc].cpu().numpy(), xout[:Sc].cpu().numpy(), vals[:Sc].cpu().numpy()
        t0 = time.perf_counter()
        W1 = oracle.weights_conservative_1d_batched(xi_h, xo_h)
        for k in range(Sc):
            oracle.regrid_from_weights(*W1[k], v_h[k:k + 1], n - 1)
        dt = time.perf_counter() - t0
        c2["cpu_baseline"] = {"value": Sc / dt, "unit": "spectra/s", "cores": oracle.num_threads(), "kind": "port",
                              "sample": f"{Sc} spectra: batched weights build + per-spectrum apply ({dt:.2f} s)"}
        got = res[:4].cpu().numpy()
        want = np.stack([oracle.regrid_from_weights(*W1[k], v_h[k:k + 1], n - 1)[0] for k in range(4)])
        c2["equals_cpu_port_bitwise_on_4_spectra"] = bool(np.array_equal(got, want))
    out["config2"] = c2
    del xin, xout, vals, res
    torch.cuda.empty_cache()

    # ---- config 4: every frame carries its own grid; 8 frames per GPU, sharded with no collective ---------
    n4, per_rank = 2049, 8
    frames = [rank * per_rank + q for q in range(per_rank)]
    grids, t_jit = [], 0.0
    for f in frames:
        gi4, go4 = cases.benchmark_family(n4, distorted=True, angle=0.4 + 0.002 * f, phase=float(f))
        t0 = time.perf_counter()
        co4 = cases.perturb_like_reference(go4, (-1, -2), 42)   # the reference's host jitter stream (NumPy)
        t_jit += time.perf_counter() - t0
        grids.append([torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (*gi4, *co4)])
    from regridding_b200 import _parallel

    # all 8 x world slices exist conceptually; this rank only materialised its own share (slice index -> position)
    class _Share:
        def __len__(self):
            return per_rank * world

        def __getitem__(self, k):
            return tuple(a.numpy() for a in grids[k - rank * per_rank])

    for _ in range(2):  # warm-up (the first call learns the buffer sizes of this shape, the second allocates them)
        _parallel.build_weights_2d_slices(_Share(), device=dev)
    barrier()
    t0 = time.perf_counter()
    built = _parallel.build_weights_2d_slices(_Share(), device=dev)
    torch.cuda.synchronize(dev)
    nnz4 = sum(d.nnz for d in built.values())
    wall = max_over_ranks(time.perf_counter() - t0)
    jit = max_over_ranks(t_jit)
    cells4 = (n4 - 1) ** 2
    out["config4"] = {"workload": f"{per_rank * world} frames x 2048^2, each with its own curvilinear grid; {per_rank} per GPU",
                      "n_gpus": world, "frames": per_rank * world, "frames_per_gpu": per_rank,
                      "device_wall_s": wall, "scope": "_parallel.build_weights_2d_slices: host coordinates -> pinned -> H2D (overlapped) -> rg_build2d_batched "
                               "(no host sync inside a chunk of 4 slices) -> device-resident triplets; no collective",
                      "Mcells_per_s": world * per_rank * cells4 / wall / 1e6, "ms_per_frame_per_gpu": wall / per_rank * 1e3,
                      "host_jitter_s_per_gpu": jit,
                      "note": "the reference's seeded jitter is a serial NumPy stream (0.14 s per 2049^2 grid): end to end it "
                              "dominates the GPU build by ~30x on any number of GPUs; see DESIGN.md",
                      "nnz_rank0": nnz4, "reference_numba_s_per_frame_8_vcpu": 72.0}
    del grids, built
    torch.cuda.empty_cache()

    # ---- config 5: cell location of 8192^2 output points in a 4096^2-vertex curvilinear grid -------------
    gi5, _ = cases.benchmark_family(4096, distorted=True)
    X, Y = T(gi5[0]), T(gi5[1])
    m = 8192
    c5 = {"workload": "find_indices of 8192^2 rectilinear output points in a 4096^2-vertex distorted curvilinear grid",
          "points": m * m, "n_gpus": world}
    for name, scale in (("inside_0.7_bbox", 0.7), ("full_bbox", 1.0)):
        cx, cy = float(X.min() + X.max()) / 2, float(Y.min() + Y.max()) / 2
        hx, hy = float(X.max() - X.min()) / 2 * scale, float(Y.max() - Y.min()) / 2 * scale
        px = torch.linspace(cx - hx, cx + hx, m, dtype=torch.float64, device=dev)[:, None].expand(m, m).contiguous()
        py = torch.linspace(cy - hy, cy + hy, m, dtype=torch.float64, device=dev)[None, :].expand(m, m).contiguous()
        ms5, idx = timed(lambda: _device.find_indices_2d(X, Y, px, py, -1), reps=5, warm=3)  # (warm: both result buffers allocated)
        c5[name] = {"ms": ms5, "Mpoints_per_s": world * m * m / ms5 / 1e3, "fraction_inside": float((idx >= 0).double().mean()),
                    "roofline": roof((16 + 8) * m * m, ms5)}
        if with_cpu and name == "full_bbox":
            from oracle import oracle

            rng = np.random.default_rng(5)
            a, b = rng.integers(0, m, 100000), rng.integers(0, m, 100000)
            pxs, pys = px[a, b].cpu().numpy(), py[a, b].cpu().numpy()
            t0 = time.perf_counter()
            want = oracle.index_of_points(gi5[0], gi5[1], pxs, pys, -1, "secant")
            dt = time.perf_counter() - t0
            c5["cpu_baseline"] = {"value": 1e5 / dt / 1e6, "unit": "Mpoints/s", "cores": oracle.num_threads(), "kind": "port",
                                  "sample": f"100000 of the points, index_of_point_secant ({dt:.2f} s)"}
            c5["equals_cpu_port_on_sample"] = bool(np.array_equal(idx[a, b].cpu().numpy(), want))
        del px, py, idx
    out["config5"] = c5
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=2049, help="vertices per axis (config 3: 2049)")
    ap.add_argument("--frames", type=int, default=1000, help="frames per GPU sharing the weights (config 3: 1000)")
    ap.add_argument("--e2e-frames", type=int, default=125)
    ap.add_argument("--cpu-frames", type=int, default=64)
    ap.add_argument("--cpu-build-n", type=int, default=1025)
    ap.add_argument("--ref-frames", type=int, default=32)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra lines for BASELINE configs 1, 2, 4, 5")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
