import sys, numpy as np
sys.path.insert(0, '.')
from kinectpy_b200 import _cabi, synth
from kinectpy_b200.pipeline import FramePipeline, PipelineConfig
from oracle import oracle as orc
sys.path.insert(0, 'tests')
import gpu_helpers as G
mode = synth.SensorMode("SMOKE", 160, 120, 126.0, 126.0, 79.5, 59.5, "hexagon")
depth, tab, T = synth.render_sequence(mode, 1, 3)
Ti = np.stack([synth.perturbed_extrinsic(T[s], 0.3, (3, -3, 3)) if s else T[s] for s in range(3)])
cfg = PipelineConfig(n_sensors=3, pixels=mode.pixels, voxel_size=0.04, sor_k=20, sor_ratio=2.0, floor_band=0.25,
                     ransac_thr=0.02, ransac_iters=256, floor_sor_k=20, floor_sor_ratio=1.0, icp_voxel=0.04,
                     icp_max_corr=0.08, normals_radius=0.08, n_streams=1)
pipe = FramePipeline(cfg, tab, T, Ti)
got = pipe.run(depth, want_points=True)[0]
ref = orc.frame_pipeline(cfg, depth[0], tab, T, Ti)
for i in range(2):
    r = ref["icp"][i]
    print("pair", i, "gpu iters", got.icp_iters[i], "fit", got.icp_fitness[i], "rmse", got.icp_rmse[i])
    print("       ora iters", r["iters"], "fit", r["fitness"], "rmse", r["rmse"], "maxdiff", np.abs(got.icp_T[i] - r["T"]).max())
# step by step with identical inputs
ctx = _cabi.default_context()
xyz, valid, _ = orc.unproject(depth[0][None], tab, T, flags=cfg.unproject_flags, scale=cfg.scale)
P = mode.pixels
tgt = orc.voxel_downsample(xyz[0][:P], cfg.icp_voxel)["points"]
nrm_o = orc.estimate_normals(tgt, cfg.normals_radius, cfg.normals_max_nn)
nrm_g = G.normals(ctx, tgt, cfg.normals_radius, cfg.normals_max_nn)
d = np.abs((nrm_o.astype(np.float64) * nrm_g).sum(1))
print("normals |dot| min", d.min(), "frac<1-1e-6", (d < 1 - 1e-6).mean(), "n", len(d))
raw, _, _ = orc.unproject(depth[0][None, 1:2], tab[1:2], None, flags=cfg.unproject_flags, scale=cfg.scale)
src = orc.voxel_downsample(raw[0], cfg.icp_voxel)["points"]
for mi in (0, 1, 2, 3, 5, 8, 30):
    a = orc.icp_point_to_plane(src, tgt, nrm_o, cfg.icp_max_corr, init=Ti[1], max_iter=mi)
    b = G.icp(ctx, src, tgt, nrm_o, cfg.icp_max_corr, init=Ti[1], max_iter=mi)
    c = G.icp(ctx, src, tgt, nrm_g, cfg.icp_max_corr, init=Ti[1], max_iter=mi)
    print("max_iter", mi, "ora", a["iters"], a["ncorr"], "%.9f" % a["rmse"], "| gpu(same normals)", b["iters"], b["ncorr"], "%.9f" % b["rmse"],
          "dT %.2e" % np.abs(a["T"] - b["T"]).max(), "| gpu(gpu normals)", c["iters"], c["ncorr"], "dT %.2e" % np.abs(a["T"] - c["T"]).max())
