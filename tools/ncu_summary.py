"""Print key metrics from an .ncu-rep (raw page)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__warps_active.avg.pct", "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed.sum.pct", "smsp__issue_active.avg.pct", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp", "smsp__warp_issue_stalled",
        "smsp__average_warps_issue_stalled", "sm__pipe_fp64_cycles_active", "smsp__inst_executed_pipe_fp64", "sm__inst_executed_pipe_lsu",
        "launch__occupancy", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "smsp__pcsamp_warps_issue_stalled",
        "smsp__warps_issue_stalled"]
for r in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in keys):
            v = r[i]
            if v in ("0", "0.000000", ""): continue
            print(f"{h:90s} {units[i]:14s} {v}")
