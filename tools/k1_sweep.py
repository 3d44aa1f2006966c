"""K1 (unproject + transform + fuse) timed alone over the batch size B: one table-stationary launch for B frames.
usage (GPU box): python tools/k1_sweep.py [WFOV|NFOV]
Algorithmic bytes per launch: B*S*P*(2 + 12) + S*P*8.  L2 flushed before every timed launch."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kinectpy_b200 import _cabi, synth

WITH_BOUNDS = os.environ.get("K1_BOUNDS", "1") == "1"
mode = synth.MODES[sys.argv[1] if len(sys.argv) > 1 else "WFOV"]
ctx = _cabi.default_context(0)
depth, tab, T = synth.render_sequence(mode, 2, 3)
T = np.ascontiguousarray(synth.scale_extrinsics(T, 1e-3), np.float64).reshape(-1)
S, P = 3, mode.pixels
d_tab = ctx.to_device(tab, np.float32)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
for B in [int(x) for x in os.environ.get("K1_BS", "1,2,4,8,12,16").split(",")]:
    batch = np.ascontiguousarray(depth[np.arange(B) % depth.shape[0]])
    d_depth = ctx.to_device(batch, np.uint16)
    xyz = ctx.empty((B, S * P, 3), np.float32)
    bounds = ctx.empty((B, 6), np.float32)
    nvalid = ctx.empty((B,), np.int32)
    def launch():
        ctx.check(ctx.lib.kp_unproject_transform(ctx.handle, d_depth.ptr, d_tab.ptr, T.ctypes.data, B, S, P, 0, 1e-3,
                                                 xyz.ptr, None, None, (bounds.ptr if WITH_BOUNDS else None), (nvalid.ptr if WITH_BOUNDS else None)))
    for _ in range(3):
        launch()
    ms = []
    for _ in range(10):
        ctx.flush_l2(); ctx.sync(); ctx.timer_start(); launch(); ms.append(ctx.timer_stop())
    t = float(np.median(ms))
    by = B * S * P * 14.0 + S * P * 8.0
    print("B=%2d  %.1f us  %.0f GB/s  frac %.3f (includes the bounds init / decode launches)" % (B, 1e3 * t, by / t / 1e6, by / t / 1e6 / peak))
