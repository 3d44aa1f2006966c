"""Whole-frame driver (BASELINE config C4 shape) against the oracle composition, stage by stage."""
import numpy as np
import pytest

from kinectpy_b200 import synth
from kinectpy_b200.pipeline import FramePipeline, PipelineConfig

pytestmark = pytest.mark.gpu


def oracle_frame(oracle, cfg, depth_f, tab, T_fuse, T_icp):
    return oracle.frame_pipeline(cfg, depth_f, tab, T_fuse, T_icp)


@pytest.mark.parametrize("scale,int16", [(1e-3, True), (1e-3, False)])
def test_pipeline_matches_oracle_composition(oracle, scale, int16):
    mode = synth.NFOV
    F = 2
    depth, tab, T = synth.render_sequence(mode, F, 3)
    T_fuse = synth.scale_extrinsics(T, scale)
    T_icp = np.stack([synth.perturbed_extrinsic(T_fuse[s], unit_scale=scale) if s else T_fuse[s] for s in range(3)])
    cfg = PipelineConfig(n_sensors=3, pixels=mode.pixels, scale=scale, n_streams=2,
                         unproject_flags=(1 if int16 else 0) | 2)
    pipe = FramePipeline(cfg, tab, T_fuse, T_icp)
    outs = pipe.run(depth, want_points=True)
    dev = pipe.upload(depth)
    outs_dev = pipe.run(dev, want_points=True)
    for f in range(F):
        ref = oracle_frame(oracle, cfg, depth[f], tab, T_fuse, T_icp)
        got = outs[f]
        assert (got.n_fused, got.n_voxel, got.n_sor, got.n_floor_inliers) == \
            (ref["n_fused"], ref["n_voxel"], ref["n_sor"], ref["n_floor_inliers"])
        assert got.n_out == len(ref["points"])
        assert np.array_equal(got.points, ref["points"])          # whole chain: identical final cloud
        assert np.array_equal(outs_dev[f].points, got.points)      # host-input and HBM-resident legs agree
        for i in range(2):
            r = ref["icp"][i]
            assert np.abs(got.icp_T[i][:3, :3] - r["T"][:3, :3]).max() < 1e-4
            assert np.abs(got.icp_T[i][:3, 3] - r["T"][:3, 3]).max() < 1e-4 * (scale / 1e-3)
            assert got.icp_fitness[i] == pytest.approx(r["fitness"], abs=2e-3)
            # the refinement pulls the perturbed start back to the ground-truth extrinsic
            err0 = np.abs(T_icp[i + 1] - T_fuse[i + 1]).max()
            err1 = np.abs(got.icp_T[i] - T_fuse[i + 1]).max()
            assert err1 < err0
    assert pipe.launch_count() > 0
    pipe.close()


def test_pipeline_frames_are_independent_of_batching(oracle):
    """Sharding contract: a frame's result does not depend on which worker / batch processed it."""
    mode = synth.NFOV
    depth, tab, T = synth.render_sequence(mode, 3, 3)
    cfg1 = PipelineConfig(n_sensors=3, pixels=mode.pixels, n_streams=1, do_icp=False)
    cfg4 = PipelineConfig(n_sensors=3, pixels=mode.pixels, n_streams=3, do_icp=False)
    a = FramePipeline(cfg1, tab, T).run(depth, want_points=True)
    b = FramePipeline(cfg4, tab, T).run(depth, want_points=True)
    c = FramePipeline(cfg1, tab, T).run(depth[1:2], want_points=True)
    for f in range(3):
        assert np.array_equal(a[f].points, b[f].points)
    assert np.array_equal(a[1].points, c[0].points)
