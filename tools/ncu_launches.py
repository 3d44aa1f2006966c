"""ncu `--metrics gpu__time_duration.sum --csv` log -> launch list of ONE frame (csv) + per-kernel summary (txt).
usage: python tools/ncu_launches.py <ncu_log.csv> <out_prefix> ["header note"]
A frame starts at the K1 launch whose grid covers all sensors (grid.y == 3) and ends before the next one."""
import csv, re, sys, collections
src, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, gi, bi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size"), hdr.index("Metric Value")
def short(n):
    n = re.sub(r"\(.*$", "", n).replace("<unnamed>::", "").replace("void ", "").strip()
    return re.sub(r"<.*$", "", n)
L = [(short(r[ki]), r[gi], r[bi], int(float(r[vi]))) for r in rows[1:]]
starts = [i for i, l in enumerate(L) if l[0] == "k_unproject" and l[1].replace(" ", "").endswith(",3,1)")]
if len(starts) >= 2:
    L = L[starts[-2]:starts[-1]]
elif starts:
    L = L[starts[-1]:]
with open(out + ".csv", "w") as f:
    f.write("kernel,grid,block,duration_ns\n")
    for l in L:
        f.write('%s,"%s","%s",%d\n' % l)
agg = collections.OrderedDict()
for k, _, _, ns in L:
    a = agg.setdefault(k, [0, 0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
with open(out + ".txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, one frame. %s\n" % note)
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py kernels{}, not absolutes\n")
    f.write("# kernel                        launches   total_us   share\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-30s %6d %10.1f %6.1f%%\n" % (k, a[0], a[1] / 1e3, 100.0 * a[1] / tot))
    f.write("total %d launches, %.1f us\n" % (len(L), tot / 1e3))
print(open(out + ".txt").read())
