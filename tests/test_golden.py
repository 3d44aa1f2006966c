"""Golden vectors produced by running the REFERENCE'S OWN numpy arithmetic (tests/golden/make_golden.py,
executed in the build container against /root/reference with a container-only open3d stand-in).
CPU part: the oracle and the host-side mirrors against the fixtures.  GPU part: the kernels, through the
reference-shaped Python surface, against the same fixtures."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_own_arithmetic.npz"))


def test_equation_plane_matches_reference(capsys):
    from kinectpy_b200.floor_removal import equation_plane
    for tri, abcd in zip(GOLD["plane_triples"], GOLD["plane_abcd"]):
        assert np.array_equal(np.array(equation_plane(tri[0], tri[1], tri[2])), abcd)
    assert "equation of plane is" in capsys.readouterr().out      # the reference prints it (floor_removal.py:35)


def test_kalman_filter_matches_reference():
    from kinectpy_b200.preprocessing.filtering import kalman_filter
    assert np.allclose(kalman_filter(GOLD["kalman_in"]), GOLD["kalman_out"], rtol=1e-12, atol=1e-12)
    assert np.allclose(kalman_filter(GOLD["kalman_in"], ri=5, qi=2, fi=0.9, hi=1), GOLD["kalman_out_params"], rtol=1e-12, atol=1e-12)


def test_oracle_plane_side_matches_reference(oracle):
    a, b, c, d = GOLD["above_plane"]
    keep = oracle.plane_side(GOLD["above_cloud"], a, b, c, d).astype(bool)
    assert np.array_equal(GOLD["above_cloud"][keep], GOLD["above_points"])


def test_oracle_xyz16_rule_and_crop_match_reference(oracle):
    pts, valid = oracle.points_from_xyz16(GOLD["dat_xyz16"])
    assert np.array_equal(pts[valid.astype(bool)].astype(np.float64), GOLD["rgbd_points"])
    keep, med = oracle.crop_mask(GOLD["crop_filtered_img"], GOLD["dat_xyz16"], 750.0)
    assert med == float(GOLD["crop_median"])
    pts, valid = oracle.points_from_xyz16(GOLD["dat_xyz16"], keep=keep)
    assert np.array_equal(pts[valid.astype(bool)].astype(np.float64), GOLD["crop_points"])


def test_oracle_band_matches_reference(oracle):
    low = oracle.band_mask(GOLD["band_cloud"], 200, 1).astype(bool)
    assert np.array_equal(np.flatnonzero(low), GOLD["band_lower_idx"])
    assert np.array_equal(np.flatnonzero(~low), GOLD["band_upper_idx"])


def test_depth_dat_reader_matches_reference(tmp_path):
    from kinectpy_b200.utils import io as kio
    fp = str(tmp_path / "99")
    GOLD["dat_xyz16"].tofile(fp + "_depth.dat")
    assert np.array_equal(kio.load_depth(fp), GOLD["dat_xyz16"])


# ------------------------------------------------------------------ GPU ----
@pytest.mark.gpu
def test_gpu_pcd_above_plane_matches_reference():
    from kinectpy_b200 import PointCloud
    from kinectpy_b200.floor_removal import pcd_above_plane
    a, b, c, d = GOLD["above_plane"]
    out = pcd_above_plane(a, b, c, d, PointCloud(GOLD["above_cloud"]))
    assert np.array_equal(np.asarray(out.points), GOLD["above_points"])


@pytest.mark.gpu
def test_gpu_rgbd_to_pointcloud_matches_reference():
    from kinectpy_b200.utils.io import rgbd_to_pointcloud
    pcd = rgbd_to_pointcloud(GOLD["dat_rgb"], GOLD["dat_xyz16"])
    assert np.array_equal(np.asarray(pcd.points), GOLD["rgbd_points"])
    assert np.allclose(np.asarray(pcd.colors), GOLD["rgbd_colors"], atol=1e-7)      # colours are float32 on the device


@pytest.mark.gpu
def test_gpu_human_crop_matches_reference(ctx):
    import gpu_helpers as G
    from kinectpy_b200.preprocessing.data import transform_filtered_image_to_pointcloud
    keep, med = G.crop_mask(ctx, GOLD["crop_filtered_img"], GOLD["dat_xyz16"], 750.0)
    assert med == float(GOLD["crop_median"])
    pcd = transform_filtered_image_to_pointcloud(GOLD["crop_filtered_img"], GOLD["dat_xyz16"])
    assert np.array_equal(np.asarray(pcd.points), GOLD["crop_points"])
    assert np.allclose(np.asarray(pcd.colors), GOLD["crop_colors"], atol=1e-7)


@pytest.mark.gpu
def test_gpu_floor_band_matches_reference():
    from kinectpy_b200 import PointCloud
    from kinectpy_b200.floor_removal import split_floor_band
    lower, upper = split_floor_band(PointCloud(GOLD["band_cloud"]), 200)
    assert np.array_equal(np.asarray(lower.points), GOLD["band_cloud"][GOLD["band_lower_idx"]])
    assert np.array_equal(np.asarray(upper.points), GOLD["band_cloud"][GOLD["band_upper_idx"]])
