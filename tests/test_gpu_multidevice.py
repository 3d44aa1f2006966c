"""Multi-GPU identity: frames are sharded round-robin over G ranks with no collective on
the frame path, so the frames a rank processes must equal, bit for bit, what one GPU produces for the same frames.
With two or more devices visible the ranks' pipelines live on different devices; on a one-GPU box both use device 0
(the sharding logic and the batch grouping still differ from the single-pipeline run).  The torchrun form of the same
check is tools/multigpu_identity.py (run with `gpurun --gpus N`, log under profiles/)."""
import numpy as np
import pytest

from kinectpy_b200 import _cabi, synth
from kinectpy_b200.pipeline import FramePipeline, PipelineConfig
from kinectpy_b200.sharding import frames_for_rank

pytestmark = pytest.mark.gpu

MODE = synth.SensorMode("SMALL", 160, 120, 126.0, 126.0, 79.5, 59.5, "hexagon")


def cfg(n_streams):
    return PipelineConfig(n_sensors=3, pixels=MODE.pixels, voxel_size=0.04, sor_k=20, sor_ratio=2.0, floor_band=0.25,
                          ransac_thr=0.02, ransac_iters=256, floor_sor_k=20, floor_sor_ratio=1.0, icp_voxel=0.04,
                          icp_max_corr=0.08, normals_radius=0.08, n_streams=n_streams)


@pytest.mark.parametrize("world", [2, 3])
def test_rank_sharded_frames_equal_the_single_gpu_results(world):
    F = 7
    depth, tab, T = synth.render_sequence(MODE, F, 3)
    Ti = np.stack([synth.perturbed_extrinsic(T[s], 0.3, (3, -3, 3)) if s else T[s] for s in range(3)])
    ndev = _cabi.load_library().kp_device_count()
    ref = FramePipeline(cfg(4), tab, T, Ti, device=0).run(depth, want_points=True)
    for rank in range(world):
        mine = frames_for_rank(F, rank, world)
        pipe = FramePipeline(cfg(2 + rank), tab, T, Ti, device=rank % ndev)
        got = pipe.run(np.ascontiguousarray(depth[mine]), want_points=True)
        for j, f in enumerate(mine):
            a, b = got[j], ref[f]
            assert (a.n_fused, a.n_voxel, a.n_sor, a.n_floor_inliers, a.n_out) == (b.n_fused, b.n_voxel, b.n_sor, b.n_floor_inliers, b.n_out)
            assert np.array_equal(a.points, b.points)
            assert np.array_equal(a.icp_T, b.icp_T) and np.array_equal(a.icp_iters, b.icp_iters)
        pipe.close()
