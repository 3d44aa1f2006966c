import sys, json
src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for ln in src:
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    print("streams", d["config"].get("streams_per_gpu"), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
          "launches", d["gpu_launches"], d["kernels_note"])
    for k, v in d["kernels"].items():
        print("    %-22s %8.4f ms/frame  calls %d" % (k, v["ms_per_frame"], v.get("calls", v.get("launch_groups", 0))))
