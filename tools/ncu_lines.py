"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: python tools_ncu_lines.py <report.ncu-rep> <kernel-name> [launch-skip] [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict()
fname = None; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; si = r.index("# Samples"); ii = r.index("Instructions Executed"); ti = r.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) <= ii: continue
    if r[0] != "": cur = (fname, r[0], r[1].strip()[:100])
    if r[2] == "": continue     # source-only row
    a = agg.setdefault(cur, [0, 0, 0])
    num = lambda x: int(x) if x.strip().lstrip('-').isdigit() else 0
    a[0] += num(r[si]); a[1] += num(r[ii]); a[2] += num(r[ti])
ts = sum(a[0] for a in agg.values()) or 1; tn = sum(a[1] for a in agg.values()) or 1
print("samples", ts, "warp-inst", tn)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst act %4.1f | %s:%s | %s" % (100 * a[0] / ts, 100 * a[1] / tn, a[2] / max(a[1], 1), k[0], k[1], k[2]))
