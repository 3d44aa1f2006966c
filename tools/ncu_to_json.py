#!/usr/bin/env python
"""profiles/r02_ncu_summary.json from the `--set full` reports of tools/prof_final.sh (read here, no GPU):
per kernel the mean over the captured launches of DRAM bytes, warp instructions, issue activity, occupancy, lanes."""
import csv, io, json, subprocess, sys, os

def rows_of(path):
    # a report (.ncu-rep: read through `ncu -i`) or the raw page already exported as CSV on the GPU box
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if len(r) >= len(hdr):
            yield idx, units, r

def num(r, idx, key):
    try:
        return float(r[idx[key]].replace(",", ""))
    except Exception:
        return None

def main(paths, out_path):
    fpl = int(os.environ.get("NCU_FRAMES_PER_LAUNCH", "0")) or None     # frames per launch of the captured command
    acc = {}
    for p in paths:
        for idx, units, r in rows_of(p):
            name = r[idx["Kernel Name"]]
            import re
            m = re.search(r"(k_[A-Za-z0-9_]+)(<[^>]*>)?", name)
            short = (m.group(1) + (m.group(2) or "")) if m else name[:40]
            bscale = lambda k: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[k]], 1.0)
            tscale = {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1e-6)
            e = acc.setdefault(short, {"launches": 0, "dram": 0.0, "inst": 0.0, "issue": 0.0, "occ": 0.0, "lanes": 0.0, "ms": 0.0, "regs": 0, "names": set(), "each": []})
            e["launches"] += 1
            e["dram"] += (num(r, idx, "dram__bytes_read.sum") or 0) * bscale("dram__bytes_read.sum") + (num(r, idx, "dram__bytes_write.sum") or 0) * bscale("dram__bytes_write.sum")
            e["inst"] += num(r, idx, "smsp__inst_executed.sum") or 0
            e["issue"] += num(r, idx, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0
            e["occ"] += num(r, idx, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0
            e["lanes"] += num(r, idx, "smsp__thread_inst_executed_per_inst_executed.ratio") or 0
            e["ms"] += (num(r, idx, "gpu__time_duration.sum") or 0) * tscale
            e["regs"] = int(num(r, idx, "launch__registers_per_thread") or 0)
            e["names"].add(name[:90])
            e["each"].append({"ms": round((num(r, idx, "gpu__time_duration.sum") or 0) * tscale, 4), "warp_instructions": int(num(r, idx, "smsp__inst_executed.sum") or 0),
                              "grid": r[idx["Grid Size"]] if "Grid Size" in idx else None})
    res = {}
    for k, e in acc.items():
        n = e["launches"]
        res[k] = {"launches": n, "dram_bytes_per_launch": round(e["dram"] / n), "warp_instructions_per_launch": round(e["inst"] / n),
                  "issue_active_frac": round(e["issue"] / n / 100.0, 4), "occupancy_frac": round(e["occ"] / n / 100.0, 4),
                  "lanes_per_instruction": round(e["lanes"] / n, 2), "ncu_ms_per_launch": round(e["ms"] / n, 4), "registers": e["regs"], "frames_per_launch": fpl, "per_launch": e["each"],
                  "source": "ncu --set full --clock-control none, raw pages %s (cold-cache, serialised: shares, not absolutes)" % ", ".join(paths),
                  "instances": sorted(e["names"])}
    json.dump(res, open(out_path, "w"), indent=1)
    for k, v in res.items():
        print(k, {a: b for a, b in v.items() if a not in ("source", "instances", "per_launch")})

if __name__ == "__main__":
    main(sys.argv[2:], sys.argv[1])
