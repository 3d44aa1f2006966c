"""Point-to-point and coloured ICP: registration_icp with
TransformationEstimationPointToPoint (manual_pointcloud_registration.py:90-98) and registration_colored_icp
(preprocessing/registration.py:89-114).  CPU tests pin the oracle against numpy; GPU tests compare the kernels
with the oracle within the north-star tolerance (1e-4 on the transform)."""
import numpy as np
import pytest

from kinectpy_b200 import synth

ICP_TOL = 1e-4


def textured_sheet(n, seed):
    """A gently curved sheet z = f(x, y) with a smooth intensity texture: geometry alone leaves the in-plane
    slide almost unconstrained, colour pins it."""
    r = np.random.default_rng(seed)
    u = r.uniform(-1, 1, (n, 2))
    z = 0.03 * np.sin(2 * u[:, 0]) + 0.02 * u[:, 1] ** 2
    pts = np.stack([u[:, 0], u[:, 1], z], axis=1)
    inten = 0.5 + 0.25 * np.sin(9 * u[:, 0]) * np.cos(7 * u[:, 1])
    col = np.stack([inten, inten, inten], axis=1)
    return pts.astype(np.float32), col.astype(np.float32)


def kabsch_numpy(s, t):
    ms, mt = s.mean(0), t.mean(0)
    U, _, Vt = np.linalg.svd((t - mt).T @ (s - ms) / len(s))
    S = np.diag([1, 1, -1.0 if np.linalg.det(U) * np.linalg.det(Vt) < 0 else 1.0])
    R = U @ S @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = mt - R @ ms
    return T


# ------------------------------------------------------------------ oracle (CPU)
def test_oracle_umeyama_vs_numpy_svd(oracle):
    r = np.random.default_rng(0)
    for trial in range(20):
        s = r.normal(size=(r.integers(3, 50), 3))
        if trial % 4 == 1:
            s[:, 2] = 0.0                                   # planar: rank-2 covariance
        if trial % 4 == 2:
            s = np.outer(r.normal(size=len(s)), [1.0, 2.0, -1.0])   # collinear: rank 1 (rotation about the line is free)
        D = synth.perturbed_extrinsic(np.eye(4), angle_deg=float(r.uniform(0, 170)), shift_mm=tuple(r.normal(0, 300, 3)), unit_scale=1e-3)
        t = s @ D[:3, :3].T + D[:3, 3] + (0 if trial % 3 else r.normal(0, 1e-3, s.shape))
        T = oracle.umeyama(s, t)
        assert abs(np.linalg.det(T[:3, :3]) - 1) < 1e-9 and np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-9)
        if trial % 4 != 2:
            assert np.abs(T - kabsch_numpy(s, t)).max() < 1e-8, trial
        else:
            assert np.abs(s @ T[:3, :3].T + T[:3, 3] - t).max() < 5e-3


def test_oracle_point_to_point_icp_recovers_planted_transform(oracle):
    from conftest import make_surface_cloud
    tgt = make_surface_cloud(6000, seed=21, outliers=0.0)
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=0.8, shift_mm=(4, -3, 5), unit_scale=1e-3)
    src = oracle.transform(tgt[::2], np.linalg.inv(D))
    res = oracle.icp_point_to_point(src, tgt, 0.05, init=np.eye(4), max_iter=50)
    assert res["fitness"] > 0.99 and np.abs(res["T"] - D).max() < 2e-3
    r0 = oracle.icp_point_to_point(src, tgt, 0.05, init=D, max_iter=0)
    assert np.array_equal(r0["T"], D) and r0["iters"] == 0


def test_oracle_point_to_point_icp_vs_independent_numpy_restatement(oracle):
    """RegistrationICP + TransformationEstimationPointToPoint(False) written out independently (cKDTree 1-NN inside
    max_corr, Kabsch by numpy SVD on the matched pairs, T <- U T, relative fitness / rmse stop): same iteration count
    and fitness as the oracle, transforms equal to rounding."""
    from conftest import make_surface_cloud
    from scipy.spatial import cKDTree
    tgt = make_surface_cloud(5000, seed=23, outliers=0.0)
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=0.6, shift_mm=(3, -4, 5), unit_scale=1e-3)
    src = oracle.transform(tgt[1::2], np.linalg.inv(D))
    max_corr, max_iter = 0.04, 40
    res = oracle.icp_point_to_point(src, tgt, max_corr, init=np.eye(4), max_iter=max_iter)
    tree = cKDTree(tgt.astype(np.float64))
    T, cur = np.eye(4), src.astype(np.float64)

    def match(p):
        d, j = tree.query(p, k=1, distance_upper_bound=max_corr)
        ok = np.isfinite(d) & (d < max_corr)
        return ok, j, ok.mean(), np.sqrt((d[ok] ** 2).mean())

    ok, j, fit, rmse = match(cur)
    iters = 0
    for it in range(max_iter):
        U = kabsch_numpy(cur[ok], tgt[j[ok]].astype(np.float64))
        T = U @ T
        cur = cur @ U[:3, :3].T + U[:3, 3]
        ok, j, nfit, nrmse = match(cur)
        iters = it + 1
        done = abs(fit - nfit) < 1e-6 and abs(rmse - nrmse) < 1e-6
        fit, rmse = nfit, nrmse
        if done:
            break
    assert res["iters"] == iters and abs(res["fitness"] - fit) < 1e-12 and abs(res["rmse"] - rmse) < 1e-9
    assert np.abs(res["T"] - T).max() < 1e-9


def test_oracle_color_gradient_of_a_linear_ramp(oracle):
    # on the plane z = 0 with intensity a x + b y the tangent-plane gradient is (a, b, 0) everywhere
    r = np.random.default_rng(3)
    p = np.zeros((4000, 3), np.float32)
    p[:, :2] = r.uniform(-1, 1, (4000, 2))
    a, b = 0.3, -0.2
    inten = 0.5 + a * p[:, 0] + b * p[:, 1]
    col = np.stack([inten] * 3, axis=1).astype(np.float32)
    nrm = np.tile(np.float32([0, 0, 1]), (4000, 1))
    it, g = oracle.color_gradient(p, col, nrm, 0.1, 30)
    inner = (np.abs(p[:, 0]) < 0.85) & (np.abs(p[:, 1]) < 0.85)
    assert np.abs(g[inner] - [a, b, 0]).max() < 2e-4          # float32 colours
    assert np.abs(it - inten).max() < 1e-6
    # fewer than 4 neighbours -> zero gradient
    far = np.float32([[10, 10, 0], [10.01, 10, 0], [10, 10.01, 0]])
    _, g2 = oracle.color_gradient(far, np.full((3, 3), 0.5, np.float32), nrm[:3], 0.1, 30)
    assert not g2.any()


def test_oracle_colored_icp_pins_the_in_plane_slide(oracle):
    tgt, tcol = textured_sheet(12000, 5)
    nrm = oracle.estimate_normals(tgt, 0.08, 30)
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=0.3, shift_mm=(12, -9, 2), unit_scale=1e-3)   # mostly a slide
    sel = np.arange(0, 12000, 2)
    src, scol = oracle.transform(tgt[sel], np.linalg.inv(D)), tcol[sel]
    res = oracle.icp_colored(src, scol, tgt, tcol, nrm, 0.05, init=np.eye(4), max_iter=60)
    assert np.abs(res["T"] - D).max() < 1.5e-3
    plane = oracle.icp_point_to_plane(src, tgt, nrm, 0.05, init=np.eye(4), max_iter=60)
    assert np.abs(res["T"] - D)[:2, 3].max() < 0.5 * np.abs(plane["T"] - D)[:2, 3].max()      # colour is what pins x, y


def test_surface_has_the_reference_estimators():
    from kinectpy_b200 import o3d
    reg = o3d.pipelines.registration
    for name in ("registration_icp", "registration_colored_icp", "TransformationEstimationPointToPoint",
                 "TransformationEstimationPointToPlane", "TransformationEstimationForColoredICP", "ICPConvergenceCriteria"):
        assert hasattr(reg, name)
    assert reg.TransformationEstimationForColoredICP().lambda_geometric == 0.968
    # compute_transformation (manual_pointcloud_registration.py:90-92) is host-side: no device needed
    from kinectpy_b200.geometry import PointCloud
    r = np.random.default_rng(1)
    s = r.normal(size=(6, 3))
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=30.0, shift_mm=(100, 50, -70), unit_scale=1e-3)
    a, b = PointCloud.__new__(PointCloud), PointCloud.__new__(PointCloud)
    class _P:            # duck-typed clouds: only .points is touched
        def __init__(self, p): self.points = p
    T = reg.TransformationEstimationPointToPoint().compute_transformation(_P(s), _P(s @ D[:3, :3].T + D[:3, 3]),
                                                                           o3d.utility.Vector2iVector([[i, i] for i in range(6)]))
    assert np.abs(T - D).max() < 1e-9


# ------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
def test_gpu_point_to_point_icp_matches_oracle(ctx, oracle):
    import gpu_helpers as G
    from conftest import make_surface_cloud
    tgt = make_surface_cloud(30000, seed=31, outliers=0.01)
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=0.7, shift_mm=(4, -3, 5), unit_scale=1e-3)
    src = oracle.transform(tgt[::2], np.linalg.inv(D))
    ref = oracle.icp_point_to_point(src, tgt, 0.05, init=np.eye(4), max_iter=30)
    got = G.icp_p2p(ctx, src, tgt, 0.05, init=np.eye(4), max_iter=30)
    assert got["iters"] == ref["iters"] and got["ncorr"] == ref["ncorr"]
    assert np.abs(got["T"] - ref["T"]).max() < ICP_TOL
    assert abs(got["fitness"] - ref["fitness"]) < 1e-12 and abs(got["rmse"] - ref["rmse"]) < 1e-9
    # few points (no re-ordering path), no overlap at all, zero iterations
    got = G.icp_p2p(ctx, src[:500], tgt, 0.05, init=np.eye(4), max_iter=10)
    ref = oracle.icp_point_to_point(src[:500], tgt, 0.05, init=np.eye(4), max_iter=10)
    assert np.abs(got["T"] - ref["T"]).max() < ICP_TOL and got["ncorr"] == ref["ncorr"]
    far = G.icp_p2p(ctx, src + np.float32(50.0), tgt, 0.05, max_iter=5)
    assert far["ncorr"] == 0 and np.array_equal(far["T"], np.eye(4))
    z = G.icp_p2p(ctx, src, tgt, 0.05, init=D, max_iter=0)
    assert np.array_equal(z["T"], D) and z["iters"] == 0


@pytest.mark.gpu
def test_gpu_color_gradient_matches_oracle(ctx, oracle):
    import gpu_helpers as G
    pts, col = textured_sheet(20000, 7)
    pts[::97] = np.nan
    nrm = oracle.estimate_normals(pts, 0.08, 30)
    _, ref = oracle.color_gradient(pts, col, nrm, 0.06, 30)
    got = G.color_gradient(ctx, pts, col, nrm, 0.06, 30)
    assert np.abs(got - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
def test_gpu_colored_icp_matches_oracle(ctx, oracle):
    import gpu_helpers as G
    tgt, tcol = textured_sheet(40000, 9)
    nrm = oracle.estimate_normals(tgt, 0.05, 30)
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=0.3, shift_mm=(8, -6, 2), unit_scale=1e-3)
    sel = np.arange(0, 40000, 2)
    src, scol = oracle.transform(tgt[sel], np.linalg.inv(D)), tcol[sel]
    ref = oracle.icp_colored(src, scol, tgt, tcol, nrm, 0.03, init=np.eye(4), max_iter=40)
    got = G.icp_colored(ctx, src, scol, tgt, tcol, nrm, 0.03, init=np.eye(4), max_iter=40)
    assert got["iters"] == ref["iters"] and got["ncorr"] == ref["ncorr"]
    assert np.abs(got["T"] - ref["T"]).max() < ICP_TOL
    assert np.abs(got["T"] - D).max() < 1.5e-3
    # lambda = 1 is point-to-plane ICP (the photometric row vanishes)
    a = G.icp_colored(ctx, src, scol, tgt, tcol, nrm, 0.03, max_iter=10, lam=1.0)
    b = G.icp(ctx, src, tgt, nrm, 0.03, max_iter=10)
    assert np.abs(a["T"] - b["T"]).max() < 1e-9 and a["ncorr"] == b["ncorr"]


@pytest.mark.gpu
def test_gpu_registration_surface_variants(oracle):
    """execute_colored_ICP_registration / registration_icp(PointToPoint) through the reference call surface."""
    from kinectpy_b200 import o3d
    from kinectpy_b200.preprocessing import registration as R
    tgt, tcol = textured_sheet(60000, 11)
    tgt_mm = tgt.astype(np.float64) * 1000.0                              # the reference works in millimetres
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=0.3, shift_mm=(8, -6, 2), unit_scale=1.0)
    master = o3d.geometry.PointCloud(); master.points = tgt_mm; master.colors = tcol.astype(np.float64)
    sub = o3d.geometry.PointCloud()
    sub.points = (tgt_mm - D[:3, 3]) @ D[:3, :3]; sub.colors = tcol.astype(np.float64)
    # registration.py:92-93: source <- master, target <- sub, so the result maps master into sub: inv(D)
    T = R.execute_colored_ICP_registration(master, sub, np.eye(4))
    assert T.shape == (4, 4) and np.abs(T - np.linalg.inv(D))[:3, :3].max() < 2e-3 and np.abs(T - np.linalg.inv(D))[:3, 3].max() < 3.0
    res = o3d.pipelines.registration.registration_icp(sub, master, 30.0, np.eye(4),
                                                      o3d.pipelines.registration.TransformationEstimationPointToPoint())
    assert res.fitness > 0.9 and res.transformation.shape == (4, 4)
