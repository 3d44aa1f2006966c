"""BASELINE full sizes: one 3-sensor WFOV 1024x1024 frame (config C4 / C3) and a 3-sensor NFOV frame (C2) through
the C ABI.  The oracle finishes such a frame in seconds on the GPU box's cores, so the whole chain is compared
with it directly; on top of that, size-independent properties are checked against numpy / scipy on the full
clouds: voxel partition and means, brute-force neighbours of sampled queries, the SOR rule recomputed from the
returned means, a recount of the RANSAC inliers, the ICP statistics re-evaluated with a KD-tree, and the
sample -> erase -> resample behaviour of the fixed-N resampler."""
import ctypes as C

import numpy as np
import pytest
from scipy.spatial import cKDTree

import gpu_helpers as G
from kinectpy_b200 import synth
from kinectpy_b200.pipeline import FramePipeline, PipelineConfig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wfov_frame():
    depth, tab, T = synth.render_sequence(synth.WFOV, 1, 3)
    T_fuse = synth.scale_extrinsics(T, 1e-3)
    T_icp = np.stack([synth.perturbed_extrinsic(T_fuse[s], 0.3, (3, -3, 3), unit_scale=1e-3) if s else T_fuse[s] for s in range(3)])
    return depth, tab, T_fuse, T_icp


def test_wfov_frame_matches_oracle_end_to_end(oracle, wfov_frame):
    depth, tab, T_fuse, T_icp = wfov_frame
    cfg = PipelineConfig(n_sensors=3, pixels=synth.WFOV.pixels, n_streams=1)
    pipe = FramePipeline(cfg, tab, T_fuse, T_icp)
    got = pipe.run(depth, want_points=True)[0]
    ref = oracle.frame_pipeline(cfg, depth[0], tab, T_fuse, T_icp)
    assert (got.n_fused, got.n_voxel, got.n_sor, got.n_floor_inliers, got.n_out) == \
        (ref["n_fused"], ref["n_voxel"], ref["n_sor"], ref["n_floor_inliers"], len(ref["points"]))
    assert got.n_fused > 2_000_000 and got.n_voxel > 900_000          # the sizes SURVEY.md 8 quotes
    assert np.array_equal(got.points, ref["points"])
    for i in range(2):
        assert np.abs(got.icp_T[i] - ref["icp"][i]["T"]).max() < 1e-4         # north-star tolerance
        # (the pass count may differ: with 490 k source points the 1e-6 fitness criterion only fires when two successive
        # passes match EXACTLY the same number of points, which the last bits of the reduction order decide)
        assert 1 <= got.icp_iters[i] <= 30 and abs(got.icp_fitness[i] - ref["icp"][i]["fitness"]) < 1e-4
    pipe.close()


def test_wfov_stage_properties(ctx, wfov_frame):
    depth, tab, T_fuse, _ = wfov_frame
    xyz = G.unproject(ctx, depth, tab, T_fuse, 3, 1e-3, want=("valid", "bounds"))
    pts = xyz["xyz"][0]
    valid = ~np.isnan(pts[:, 0])
    assert valid.sum() == int(xyz["nvalid"][0]) and np.array_equal(valid, xyz["valid"][0].astype(bool))
    # ---- voxel: a partition of the valid points; means are the per-voxel averages
    v = G.voxel(ctx, pts, 0.01)
    m, pv = v["m"], v["point_voxel"]
    assert np.array_equal(pv >= 0, valid) and pv.max() == m - 1 and len(np.unique(pv[valid])) == m
    ijk = v["ijk"]
    key = (ijk[:, 0].astype(np.int64) << 42) | (ijk[:, 1].astype(np.int64) << 21) | ijk[:, 2].astype(np.int64)
    assert np.all(np.diff(key) > 0)                                   # sorted by (ix, iy, iz), unique
    P = pts[valid].astype(np.float64)
    cnt = np.bincount(pv[valid], minlength=m)
    mean = np.stack([np.bincount(pv[valid], P[:, c], m) for c in range(3)], 1) / cnt[:, None]
    assert np.abs(mean - v["points"]).max() < 1e-5                    # north-star coordinate tolerance
    exp_ijk = np.floor((P - v["min_bound"]) / 0.01).astype(np.int32)
    assert np.array_equal(exp_ijk, ijk[pv[valid]])                    # voxel keys, all 2.4 M of them
    # ---- neighbours: sampled queries against a brute-force scan of the whole cloud
    cloud = v["points"]
    k = 20
    idx, d2, cntk = G.knn(ctx, cloud, k)
    r = np.random.default_rng(0)
    C64 = cloud.astype(np.float64)
    for q in r.integers(0, m, 40):
        dx, dy, dz = C64[q, 0] - C64[:, 0], C64[q, 1] - C64[:, 1], C64[q, 2] - C64[:, 2]
        dd = (dx * dx + dy * dy) + dz * dz
        order = np.lexsort((np.arange(m), dd))[:k]
        assert np.array_equal(idx[q], order) and np.array_equal(d2[q], dd[order])
    assert np.all(idx[:, 0] == np.arange(m)) and np.all(np.diff(d2, axis=1) >= 0) and np.all(cntk == k)
    tree = cKDTree(C64)
    dk, _ = tree.query(C64[::97], k=k)
    assert np.allclose(np.sqrt(d2[::97]), dk, rtol=0, atol=1e-12)
    # ---- SOR: the rule recomputed from the returned means
    keep, mean_d, stats, kept = G.sor(ctx, cloud, k, 2.0, 0.01 * 1.5 * np.sqrt(k / np.pi))
    assert np.allclose(mean_d, np.sqrt(d2).sum(1) / k, rtol=1e-14)
    mu = mean_d[mean_d > 0].sum() / m
    sd = np.sqrt(((mean_d[mean_d > 0] - mu) ** 2).sum() / (m - 1))
    assert abs(stats[0] - mu) < 1e-12 and abs(stats[1] - sd) < 1e-12
    thr = stats[2]
    assert np.array_equal(keep.astype(bool), (mean_d > 0) & (mean_d < thr)) and kept == keep.sum()
    # ---- RANSAC (config C3: 1 cm, 1000 hypotheses) on the fused cloud: recount the winner's inliers
    plane, mask, best, counts, ninl = G.ransac(ctx, cloud, 0.01, 3, 1000)
    assert counts[best] == counts.max() == ninl == mask.sum()
    ids = [int(synth_rng(1234, best, j) % m) for j in range(3)]
    p0, p1, p2 = C64[ids]
    nrm = np.cross(p1 - p0, p2 - p0)
    nrm /= np.linalg.norm(nrm)
    dist = np.abs(C64 @ nrm - nrm @ p0)
    assert abs(int((dist < 0.01).sum()) - int(ninl)) <= 2             # numpy's dot order differs in the last bit
    assert abs(abs(plane[:3] @ nrm) - 1) < 1e-3                        # the refit plane is the same plane


def synth_rng(seed, a, b):
    def mix(z):
        z = (z + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    return mix((mix((mix(seed) + a) & 0xFFFFFFFFFFFFFFFF) + b) & 0xFFFFFFFFFFFFFFFF)


def test_nfov_icp_pair_statistics(ctx):
    """Config C2: the converged transform's fitness / rmse re-evaluated with a KD-tree."""
    depth, tab, T = synth.render_sequence(synth.NFOV, 1, 3)
    T_fuse = synth.scale_extrinsics(T, 1e-3)
    P = synth.NFOV.pixels
    tgt_raw = G.unproject(ctx, depth[:, :1], tab[:1], T_fuse[:1], 3, 1e-3, want=())["xyz"][0]
    src_raw = G.unproject(ctx, depth[:, 1:2], tab[1:2], None, 3, 1e-3, want=())["xyz"][0]
    tgt = G.voxel(ctx, tgt_raw, 0.01)["points"]
    src = G.voxel(ctx, src_raw, 0.01)["points"]
    nrm = G.normals(ctx, tgt, 0.02, 30)
    init = synth.perturbed_extrinsic(T_fuse[1], 0.3, (3, -3, 3), unit_scale=1e-3)
    res = G.icp(ctx, src, tgt, nrm, 0.02, init=init, max_iter=30)
    assert 1 <= res["iters"] <= 30 and np.abs(res["T"] - T_fuse[1]).max() < np.abs(init - T_fuse[1]).max()
    moved = src.astype(np.float64) @ res["T"][:3, :3].T + res["T"][:3, 3]
    d, _ = cKDTree(tgt.astype(np.float64)).query(moved, k=1)
    inl = d < 0.02
    assert abs(inl.mean() - res["fitness"]) < 2e-4 and abs(np.sqrt((d[inl] ** 2).mean()) - res["rmse"]) < 1e-6
    assert abs(int(inl.sum()) - res["ncorr"]) <= 0.0002 * len(src)


def test_resample_sample_erase_resample(ctx):
    """K6 at the C5 size: the sample is a subset without repeats; erasing the sampled rows (NaN) and sampling
    again with the same stream picks only rows that were not picked before."""
    r = np.random.default_rng(5)
    cloud = (r.random((640000, 3)) * 4).astype(np.float32)
    p1, i1 = G.resample(ctx, cloud, 4096, 0, 42, 7)
    assert len(set(i1.tolist())) == 4096 and np.array_equal(p1, cloud[i1])
    erased = cloud.copy()
    erased[i1] = np.nan
    p2, i2 = G.resample(ctx, erased, 4096, 0, 42, 7)
    assert not set(i1.tolist()) & set(i2.tolist()) and not np.isnan(p2).any()
    # the keys are a property of (seed, stream, index): the second sample is the next 4096 in key order
    both, ib = G.resample(ctx, cloud, 8192, 0, 42, 7)
    assert np.array_equal(ib[:4096], i1) and np.array_equal(ib[4096:], i2)
