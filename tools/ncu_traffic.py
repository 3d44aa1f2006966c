"""ncu report -> per-kernel DRAM traffic summary (JSON) that bench.py reads for `roofline.traffic`.
usage: python tools/ncu_traffic.py <report.ncu-rep> [more reports ...] > profiles/rNN_traffic.json
Per kernel name: launches seen, mean duration (us, cold-cache under ncu), mean dram read+write bytes per launch."""
import csv, json, re, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
out = {}
for rep in sys.argv[1:]:
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        continue
    h, u = rows[0], rows[1]
    col = {name: i for i, name in enumerate(h)}
    for r in rows[2:]:
        name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", "").strip()
        rd = float(r[col["dram__bytes_read.sum"]]) * UNIT[u[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * UNIT[u[col["dram__bytes_write.sum"]]]
        dur = float(r[col["gpu__time_duration.sum"]]) * TIME[u[col["gpu__time_duration.sum"]]]
        e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "us": 0.0, "report": rep.split("/")[-1]})
        e["launches"] += 1; e["dram_bytes"] += rd + wr; e["us"] += dur
for e in out.values():
    e["dram_bytes_per_launch"] = e.pop("dram_bytes") / e["launches"]
    e["us_per_launch_under_ncu"] = round(e.pop("us") / e["launches"], 2)
json.dump(out, sys.stdout, indent=1, sort_keys=True)
print()
