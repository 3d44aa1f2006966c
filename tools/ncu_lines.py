#!/usr/bin/env python
"""Per-CUDA-source-line totals (instructions executed, stall samples) from `ncu --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None
hdr = None
agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        i_thr = hdr.index("Thread Instructions Executed")
        continue
    if hdr and r and r[0] not in ("", "Function Name") and r[2] == "-":
        try:
            agg.append((int(r[i_inst]), int(r[i_samp]), int(r[i_thr]), cur_file, r[0], r[1][:110]))
        except (ValueError, IndexError):
            pass
tot_i = sum(a[0] for a in agg) or 1
tot_s = sum(a[1] for a in agg) or 1
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
print("--- by instructions")
for a in sorted(agg, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% samp  lanes %4.1f  %s:%s  %s" % (100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s, a[2] / max(a[0], 1), a[3], a[4], a[5]))
print("--- by stall samples")
for a in sorted(agg, key=lambda x: -x[1])[:top]:
    print("%5.1f%% inst %5.1f%% samp  lanes %4.1f  %s:%s  %s" % (100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s, a[2] / max(a[0], 1), a[3], a[4], a[5]))
