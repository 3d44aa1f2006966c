#!/usr/bin/env python
"""Condensed per-launch summary of an ncu report (read here, no GPU): duration, instructions, issue activity,
occupancy, lanes per instruction, L1 / L2 hit rates, DRAM bytes, top stall reasons."""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("smsp__inst_executed.sum", "warp_inst_M", 1e-6),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_inst", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct", 1),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct", 1),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct", 1),
    ("launch__grid_size", "grid", 1),
]
STALL = "smsp__average_warp_latency_issue_stalled_"  # (raw page name prefix varies; fall back to pcsamp)

def main(path):
    # a report (.ncu-rep: read through `ncu -i`) or the raw page already exported as CSV on the GPU box
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    tscale = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}
    name_i = idx.get("Kernel Name")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        vals = {}
        for k, short, sc in KEYS:
            if k in idx:
                try:
                    if short == "dur_us":
                        sc = tscale.get(units[idx[k]], 1e-3)
                    elif short.endswith("_MB"):
                        sc = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[idx[k]], 1e-6)
                    vals[short] = round(float(r[idx[k]].replace(",", "")) * sc, 3)
                except ValueError:
                    vals[short] = r[idx[k]]
        stalls = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print(r[name_i][:70], "|", " ".join("%s=%s" % kv for kv in vals.items()))
        print("      stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6]))

if __name__ == "__main__":
    main(sys.argv[1])
