"""The reference-shaped Python surface (same names / defaults as preprocessing/filtering.py,
preprocessing/registration.py, floor_removal.py, utils/io.py) against the oracle compositions."""
import copy

import numpy as np
import pytest

from conftest import make_surface_cloud
from kinectpy_b200 import PointCloud, KinectPyB200Error, synth
from kinectpy_b200 import floor_removal as fr
from kinectpy_b200 import o3d
from kinectpy_b200.preprocessing import filtering, registration
from kinectpy_b200.utils import io as kio

pytestmark = pytest.mark.gpu


def f32(a):
    return np.asarray(a, dtype=np.float32)


def test_filter_outliers_matches_oracle(oracle):
    pts = make_surface_cloud(40000, seed=1, scale=1000.0)            # millimetres, like the reference
    pcd = PointCloud(pts.astype(np.float64))
    pcd.colors = np.random.default_rng(0).random((len(pts), 3))
    before = np.asarray(pcd.points).copy()
    out = filtering.filter_outliers(pcd, nb_neighbors=20, std_ratio=2.0, voxel_size=35)
    ref = oracle.filter_outliers(pts, 20, 2.0, 35)
    assert np.array_equal(f32(out.points), ref)
    assert out.has_colors() and len(out.colors) == len(ref)
    assert np.array_equal(np.asarray(pcd.points), before)            # input untouched (deepcopy in the reference)
    # reference defaults: k=200, 3 sigma, voxel 0.02 on mm data only merges duplicates
    small = make_surface_cloud(3000, seed=2, scale=1000.0)
    assert np.array_equal(f32(filtering.filter_outliers(PointCloud(small)).points), oracle.filter_outliers(small))


def test_pointcloud_surface(oracle):
    pts = make_surface_cloud(5000, seed=3)
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(pts.astype(np.float64))
    assert len(pcd) == 5000 and pcd.has_points() and not pcd.has_normals() and not pcd.has_colors()
    assert np.array_equal(pcd.get_min_bound(), pts.min(0).astype(np.float64))
    assert np.array_equal(pcd.get_max_bound(), pts.max(0).astype(np.float64))
    cp = copy.deepcopy(pcd)
    T = synth.extrinsics(3)[1]
    pcd.transform(T)
    assert np.array_equal(f32(pcd.points), oracle.transform(pts, T))
    assert np.array_equal(f32(cp.points), pts)
    sel = pcd.select_by_index(np.argwhere(pts[:, 1] > 0))           # argwhere -> (n,1), as floor_removal.py:65 passes it
    assert np.array_equal(f32(sel.points), oracle.transform(pts, T)[pts[:, 1] > 0])
    inv = pcd.select_by_index(np.argwhere(pts[:, 1] > 0), invert=True)
    assert len(sel) + len(inv) == len(pcd)
    both = sel + inv
    assert np.array_equal(f32(both.points), np.concatenate([f32(sel.points), f32(inv.points)]))
    down, ijk, pv = cp.voxel_down_sample_and_trace(0.05)
    ref = oracle.voxel_downsample(pts, 0.05)
    assert np.array_equal(ijk, ref["ijk"]) and np.array_equal(pv, ref["point_voxel"])
    cl, ind = cp.remove_statistical_outlier(20, 2.0)
    keep, _, _ = oracle.sor(pts, 20, 2.0)
    assert ind == np.flatnonzero(keep).tolist() and np.array_equal(f32(cl.points), pts[keep.astype(bool)])
    cl, ind = cp.remove_radius_outlier(5, 0.05)
    keep, _ = oracle.radius_outlier(pts, 5, 0.05)
    assert ind == np.flatnonzero(keep).tolist()
    cp.estimate_normals(o3d.geometry.KDTreeSearchParamHybrid(radius=0.1, max_nn=30))
    assert cp.has_normals() and np.allclose(np.linalg.norm(cp.normals, axis=1), 1, atol=1e-5)
    with pytest.raises(KinectPyB200Error):
        cp.voxel_down_sample(0)
    with pytest.raises(KinectPyB200Error):
        cp.remove_statistical_outlier(0, 1.0)
    empty = PointCloud()
    assert len(empty.voxel_down_sample(1.0)) == 0 and empty.remove_statistical_outlier(5, 1.0)[1] == []


def test_floor_removal_matches_oracle(oracle):
    r = np.random.default_rng(4)
    n = 30000
    room = np.stack([r.uniform(-2000, 2000, n), r.uniform(-1500, 1200, n), r.uniform(500, 4000, n)], 1)
    room[: n // 2, 1] = 1200 + r.normal(0, 3, n // 2)                 # floor at max-y (y points down), millimetres
    room = room.astype(np.float32)
    pcd = PointCloud(room)
    out, plane, inliers = fr.remove_floor(pcd, seed=1234, return_details=True)   # reference literals: 200/30/30/2000/50/0.30
    ref_pts, ref_plane, ref_inl = oracle.remove_floor(room, seed=1234)
    assert inliers == np.flatnonzero(ref_inl).tolist()                # RANSAC inlier set: bit-exact
    assert np.allclose(plane, ref_plane, atol=1e-9)
    assert np.array_equal(f32(out.points), ref_pts)
    assert abs(abs(plane[1]) - 1) < 1e-3
    # equation_plane / pcd_above_plane (floor_removal.py:21-51)
    a, b, c, d = fr.equation_plane((0, 1200, 0), (1, 1200, 0), (0, 1200, 1))
    above = fr.pcd_above_plane(a, b, c, d, pcd)
    keep = oracle.plane_side(room, a, b, c, d)
    assert np.array_equal(f32(above.points), room[keep.astype(bool)])


def test_point_to_plane_registration_matches_oracle(oracle):
    master = make_surface_cloud(30000, seed=5, outliers=0.0, scale=1000.0)
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=1.0, shift_mm=(5, -5, 5), unit_scale=1.0)
    sub = oracle.transform(make_surface_cloud(30000, seed=6, outliers=0.0, scale=1000.0), np.linalg.inv(D))
    res = registration.execute_point_to_plane_registration(PointCloud(master), PointCloud(sub), np.eye(4), voxel_size=35,
                                                           return_result=True)
    # oracle composition of registration.py:65-86: source = sub down, target = master down (+normals, r=70, nn=40)
    tgt = oracle.voxel_downsample(master, 35)["points"]
    src = oracle.voxel_downsample(sub, 35)["points"]
    nrm = oracle.estimate_normals(tgt, 70, 40)
    ref = oracle.icp_point_to_plane(src, tgt, nrm, 100, init=np.eye(4), max_iter=30)
    T = res.transformation
    assert np.abs(T[:3, :3] - ref["T"][:3, :3]).max() < 1e-4 and np.abs(T[:3, 3] - ref["T"][:3, 3]).max() < 1e-4 * 1000
    assert np.abs(T[:3, :3] - D[:3, :3]).max() < 5e-3 and np.abs(T[:3, 3] - D[:3, 3]).max() < 5.0
    plain = registration.execute_point_to_plane_registration(PointCloud(master), PointCloud(sub), np.eye(4))
    assert np.array_equal(plain, T)


def test_rgbd_to_pointcloud_and_formats(oracle, tmp_path):
    depth, tab, T = synth.render_sequence(synth.NFOV, 1, 1)
    _, _, x16 = oracle.unproject(depth, tab, None, flags=1, scale=1.0, want_xyz16=True)
    x16 = x16[0]
    rgb = np.random.default_rng(7).integers(0, 256, (x16.shape[0], 3)).astype(np.uint8)
    fp = str(tmp_path / "1234")
    kio.save_depth(fp, x16)
    loaded = kio.load_depth(fp)
    assert loaded.dtype == np.int16 and np.array_equal(loaded, x16)
    pcd = kio.rgbd_to_pointcloud(rgb, loaded)
    ok = (x16 != 0).all(1)                                           # utils/io.py:36
    assert np.array_equal(np.asarray(pcd.points), x16[ok].astype(np.float64))
    assert np.allclose(np.asarray(pcd.colors), rgb[ok] / 255.0, atol=1e-7)
    out = str(tmp_path / "cloud.pcd")
    o3d.io.write_point_cloud(out, pcd)
    back = o3d.io.read_point_cloud(out)
    assert np.array_equal(f32(back.points), f32(pcd.points)) and back.has_colors()
    assert np.abs(np.asarray(back.colors) - np.asarray(pcd.colors)).max() <= 1 / 255 + 1e-6


def test_select_by_index_keeps_the_given_order_and_non_finite_rows_are_absent():
    """(i) an unsorted / repeated index list is gathered in the order given, like Open3D (ascending unique
    lists -- what the reference passes -- go through the mask + ordered compaction path); (ii) a row with ANY non-finite
    component is absent on the device and dropped by remove_non_finite_points; (iii) rgbd_to_pointcloud refuses depth
    values that are not int16 instead of wrapping them."""
    from kinectpy_b200 import _cabi
    from kinectpy_b200.geometry import PointCloud
    from kinectpy_b200.utils.io import rgbd_to_pointcloud
    pts = np.random.default_rng(5).random((100, 3))
    pc = PointCloud(pts)
    idx = [7, 3, 3, 50, 0]
    assert np.allclose(np.asarray(pc.select_by_index(idx).points), pts[idx].astype(np.float32))
    asc = [0, 3, 7, 50]
    assert np.allclose(np.asarray(pc.select_by_index(asc).points), pts[asc].astype(np.float32))
    assert len(pc.select_by_index(idx, invert=True).points) == 100 - 4
    bad = pts.copy()
    bad[3, 1] = np.nan
    bad[10, 2] = np.inf
    bad[20, 0] = -np.inf
    pb = PointCloud(bad)
    down = pb.voxel_down_sample(1e-3)                     # every finite point its own voxel
    assert len(down.points) == 97 and np.isfinite(np.asarray(down.points)).all()
    assert len(PointCloud(bad).remove_non_finite_points().points) == 97
    with pytest.raises(_cabi.KinectPyB200Error):
        rgbd_to_pointcloud(None, np.array([[1.0, 2.0, 70000.0]]))
    ok = rgbd_to_pointcloud(None, np.array([[1.0, 2.0, 300.0], [0.0, 5.0, 6.0]]))
    assert len(ok.points) == 1
