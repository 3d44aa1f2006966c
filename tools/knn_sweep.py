"""GPU experiment: SOR neighbour-search cost on the bench's voxelised WFOV cloud for several grid/cascade settings.
Each setting runs in a fresh process (the knobs are read once)."""
import os, subprocess, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))

def child():
    import ctypes as C
    from kinectpy_b200 import _cabi, synth
    import gpu_helpers as G
    ctx = _cabi.default_context()
    cache = '/tmp/kp_sweep_cloud.npy'
    if os.path.exists(cache):
        pts = np.load(cache)
    else:
        depth, tab, T = synth.render_sequence(synth.WFOV, 1, 3)
        xyz = G.unproject(ctx, depth, tab, T, 3, 1e-3, want=())["xyz"][0]
        pts = G.voxel(ctx, xyz, 0.01)["points"]
        np.save(cache, pts)
    k = int(os.environ.get("SW_K", "20")); mult = float(os.environ.get("SW_BASE", "1.1"))
    hint = 0.01 * mult * float(np.sqrt(k / np.pi))
    d = ctx.to_device(pts, np.float32)
    n = d.shape[0]
    keep = ctx.empty((n,), np.uint8)
    kept = C.c_int64()
    for it in range(3):
        if it == 1:
            ctx.profile(True)
        ctx.check(ctx.lib.kp_sor_mask(ctx.handle, d.ptr, n, k, 2.0, hint, keep.ptr, None, None, C.byref(kept)))
    pr = ctx.profile_read()
    out = {kk: round(v["ms"] / 2, 3) for kk, v in pr.items() if kk.startswith("knn") or kk in ("sor_knn", "radix_sort", "grid_hash")}
    print(json.dumps({"k": k, "base": mult, "slack": os.environ.get("KP_KNN_SLACK"), "kept": kept.value, "n": n, "ms": out}))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        cfgs = []
        for k in (20, 50):
            for slack in (8,):
                cfgs.append(dict(SW_K=k, SW_BASE=1.5, KP_KNN_SLACK=slack))
        for c in cfgs:
            env = dict(os.environ, KP_DEBUG_KNN="1", **{a: str(b) for a, b in c.items()})
            r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            lev = [l.split("[kp knn] ")[1] for l in r.stderr.splitlines() if "kp knn" in l][-3:]
            print(r.stdout.strip(), "|", "; ".join(dict.fromkeys(lev)))
            if r.returncode != 0:
                print(r.stderr[-2000:])
