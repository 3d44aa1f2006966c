"""`ncu --set full` report -> short text summary (first launch of each kernel), the format of profiles/*_ncu_full_summary.txt.
usage: python tools/ncu_summary.py <report.ncu-rep> "header note" > profiles/rNN_x_ncu_full_summary.txt"""
import csv, re, subprocess, sys
reps = [a for a in sys.argv[1:] if a.endswith(".csv") or a.endswith(".ncu-rep")]
note = " ".join(a for a in sys.argv[1:] if a not in reps)
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
print("# ncu --set full --clock-control none.  %s" % note)
print("# raw reports / `--page raw --csv` exports: %s (scratch, not committed); longest launch of each kernel shown\n" % ", ".join(r.split("/")[-1] for r in reps))
seen = set()
allrows = []
def _dur(r, col):
    try: return float(r[col["gpu__time_duration.sum"]])
    except Exception: return 0.0
for rep in reps:
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    allrows += [(r, col, u) for r in rows[2:]]
# the longest launch of each kernel name (an ICP pass launched after convergence returns at once)
best = {}
for r, col, u in allrows:
    nm = r[col["Kernel Name"]] + r[col["launch__grid_size"]]
    if nm not in best or _dur(r, col) > _dur(best[nm][0], best[nm][1]) * (1000.0 if u[col["gpu__time_duration.sum"]] != best[nm][2][best[nm][1]["gpu__time_duration.sum"]] else 1.0):
        best[nm] = (r, col, u)
for r, col, u in best.values():
    name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", "").strip()
    name += " grid=%s" % r[col["launch__grid_size"]] if name.startswith("k_knn_hist") else ""
    if name in seen:
        continue
    seen.add(name)
    print("[%s]" % name)
    for w in WANT:
        if w in col:
            print("  %-82s %s %s" % (w, r[col[w]], u[col[w]]))
    print()
