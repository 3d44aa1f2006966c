"""The oracle itself, pinned against independent implementations (numpy brute force,
scipy.spatial.cKDTree, numpy.linalg) -- the reference ships no golden vectors."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from conftest import make_surface_cloud


def brute_knn(pts, q, k, r2=None):
    P = pts.astype(np.float64)
    out_i, out_d = [], []
    for x in q.astype(np.float64):
        dx, dy, dz = x[0] - P[:, 0], x[1] - P[:, 1], x[2] - P[:, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        idx = np.lexsort((np.arange(len(P)), d2))
        if r2 is not None:
            idx = idx[d2[idx] < r2]
        idx = idx[:k]
        out_i.append(np.pad(idx, (0, k - len(idx)), constant_values=-1))
        out_d.append(np.pad(d2[idx], (0, k - len(idx)), constant_values=np.inf))
    return np.array(out_i, dtype=np.int32), np.array(out_d)


def test_rng_matches_numpy(oracle):
    from kinectpy_b200 import synth
    a = np.arange(50, dtype=np.uint64)
    got = np.array([oracle.rng(1234, int(x), 7) for x in a], dtype=np.uint64)
    assert np.array_equal(got, synth.rng_u64(1234, a, np.uint64(7)))


def test_csum_shape(oracle):
    r = np.random.default_rng(1)
    for n in (0, 1, 31, 32, 33, 1024, 1025, 5000, 1024 * 1024 + 3):
        x = r.normal(size=n) * 10.0 ** r.integers(-3, 3, n)
        # independent numpy restatement of the 32-lane / butterfly tree
        cur = x.copy()
        while True:
            m = (len(cur) + 1023) // 1024 if len(cur) else 1
            pad = np.zeros(m * 1024)
            pad[:len(cur)] = cur
            g = pad.reshape(m, 32, 32)
            v = np.zeros((m, 32))
            for rr in range(32):
                v = v + g[:, rr, :]
            for s in (16, 8, 4, 2, 1):
                v = v + v[:, np.arange(32) ^ s]
            cur = v[:, 0]
            if m == 1:
                break
        assert oracle.csum(x) == cur[0]
        assert oracle.csum(x) == pytest.approx(np.sum(x), rel=1e-10, abs=1e-9)


@pytest.mark.parametrize("n,k", [(300, 1), (2000, 20), (1500, 50), (100, 200)])
def test_knn_vs_bruteforce(oracle, n, k):
    pts = make_surface_cloud(n, seed=n, quantum=0.01)   # quantised: plenty of exact distance ties
    idx, d2, cnt = oracle.knn(pts, k)
    bi, bd = brute_knn(pts, pts, k)
    assert np.array_equal(idx, bi)
    assert np.array_equal(d2, bd)
    assert np.all(cnt == min(k, n))


def test_knn_hybrid_and_external_queries(oracle):
    pts = make_surface_cloud(1500, seed=3)
    q = make_surface_cloud(200, seed=4) + 0.01
    idx, d2, cnt = oracle.knn(pts, 30, queries=q, radius=0.08)
    bi, bd = brute_knn(pts, q, 30, r2=0.08 * 0.08)
    assert np.array_equal(idx, bi) and np.array_equal(d2, bd)
    assert np.array_equal(cnt, (bi >= 0).sum(1))


def test_knn_distances_vs_ckdtree(oracle):
    pts = make_surface_cloud(5000, seed=5)
    idx, d2, _ = oracle.knn(pts, 20)
    dd, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=20)
    assert np.allclose(np.sqrt(d2), dd, rtol=1e-12, atol=1e-15)


def test_voxel_vs_numpy_unique(oracle):
    pts = make_surface_cloud(20000, seed=6)
    v = 0.05
    out = oracle.voxel_downsample(pts, v)
    P = pts.astype(np.float64)
    minb = P.min(0) - v * 0.5
    ijk = np.floor((P - minb) / v).astype(np.int64)
    uniq, inv = np.unique(ijk, axis=0, return_inverse=True)
    assert out["m"] == len(uniq)
    assert np.array_equal(out["ijk"], uniq.astype(np.int32))          # np.unique sorts rows lexicographically
    assert np.array_equal(out["point_voxel"], inv.reshape(-1).astype(np.int32))
    sums = np.zeros((len(uniq), 3))
    np.add.at(sums, inv.reshape(-1), P)
    cnt = np.bincount(inv.reshape(-1), minlength=len(uniq))[:, None]
    assert np.allclose(out["points"], sums / cnt, atol=1e-6)
    # idempotence: voxel means stay inside their voxel, so a second pass at the same grid keeps the count
    again = oracle.voxel_downsample(out["points"], v)
    assert again["m"] <= out["m"]


def test_voxel_nan_and_errors(oracle):
    pts = make_surface_cloud(1000, seed=7)
    pts[::7] = np.nan
    out = oracle.voxel_downsample(pts, 0.1)
    ref = oracle.voxel_downsample(pts[~np.isnan(pts[:, 0])], 0.1)
    assert np.array_equal(out["points"], ref["points"])
    assert np.all(out["point_voxel"][::7] == -1)
    with pytest.raises(ValueError):
        oracle.voxel_downsample(pts, 0.0)
    assert oracle.voxel_downsample(np.zeros((0, 3), np.float32), 0.1)["m"] == 0


def test_sor_restates_open3d_rule(oracle):
    pts = make_surface_cloud(4000, seed=8, outliers=0.02)
    k, ratio = 20, 2.0
    keep, mean, stats = oracle.sor(pts, k, ratio)
    _, bd = brute_knn(pts, pts, k)
    m = np.sqrt(bd).mean(1)
    assert np.allclose(mean, m, rtol=1e-13)
    mu = m[m > 0].sum() / len(m)
    sd = np.sqrt((np.where(m > 0, (m - mu) ** 2, 0)).sum() / (len(m) - 1))
    assert stats[0] == pytest.approx(mu, rel=1e-12) and stats[1] == pytest.approx(sd, rel=1e-12)
    expect = (m > 0) & (m < mu + ratio * sd)
    assert (keep.astype(bool) != expect).sum() == 0
    assert 0 < keep.sum() < len(keep)
    # duplicates only: every mean is 0 -> everything is dropped
    dup = np.tile(pts[:1], (50, 1))
    assert oracle.sor(dup, 5, 1.0)[0].sum() == 0
    with pytest.raises(ValueError):
        oracle.sor(pts, 0, 1.0)


def test_radius_outlier_strict(oracle):
    pts = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [10, 0, 0]], dtype=np.float32)
    keep, cnt = oracle.radius_outlier(pts, 1, 1.0)       # d == r is NOT a neighbour (strict)
    assert cnt.tolist() == [1, 1, 1, 1] and keep.sum() == 0
    keep, cnt = oracle.radius_outlier(pts, 1, 1.0001)
    assert cnt.tolist() == [2, 3, 2, 1] and keep.tolist() == [1, 1, 1, 0]
    big = make_surface_cloud(3000, seed=9)
    keep, cnt = oracle.radius_outlier(big, 5, 0.05)
    tree = cKDTree(big.astype(np.float64))
    ref = np.array([len(x) for x in tree.query_ball_point(big.astype(np.float64), 0.05 * (1 - 1e-12))])
    assert np.array_equal(cnt, ref)


def test_eigvec_vs_numpy(oracle):
    r = np.random.default_rng(10)
    for _ in range(200):
        A = r.normal(size=(3, 3)) * r.uniform(0.01, 1, 3)
        S = A @ A.T
        n = oracle.smallest_eigvec([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]])
        w, V = np.linalg.eigh(S)
        if w[1] - w[0] < 1e-6 * w[2]:
            continue
        assert abs(abs(n @ V[:, 0]) - 1.0) < 1e-8
    assert oracle.smallest_eigvec([2.0, 0, 0, 1.0, 0, 3.0]).tolist() == [0, 1, 0]
    assert oracle.smallest_eigvec([0.0] * 6).tolist() == [0, 0, 0]


def test_normals_on_plane(oracle):
    r = np.random.default_rng(11)
    uv = r.uniform(-1, 1, (3000, 2))
    nrm = np.array([0.3, -0.5, 0.8]); nrm /= np.linalg.norm(nrm)
    e1 = np.cross(nrm, [1, 0, 0]); e1 /= np.linalg.norm(e1)
    e2 = np.cross(nrm, e1)
    pts = (uv[:, :1] * e1 + uv[:, 1:] * e2).astype(np.float32)
    out = oracle.estimate_normals(pts, 0.2, 30)
    assert np.all(np.abs(np.abs(out @ nrm) - 1) < 1e-4)
    lonely = np.array([[0, 0, 0], [5, 5, 5]], dtype=np.float32)
    assert oracle.estimate_normals(lonely, 0.1, 30).tolist() == [[0, 0, 1], [0, 0, 1]]


def test_normals_vs_independent_hybrid_search(oracle):
    """estimate_normals(KDTreeSearchParamHybrid(r, max_nn)) written out independently: the max_nn nearest points
    strictly inside r (the point itself included), covariance from the cumulants, smallest-eigenvalue eigenvector
    by numpy.linalg.eigh; fewer than 3 neighbours -> (0, 0, 1).  Compared modulo sign where the eigen-gap is clear."""
    pts = make_surface_cloud(3000, seed=31, outliers=0.01)
    radius, max_nn = 0.06, 30
    out = oracle.estimate_normals(pts, radius, max_nn)
    P = pts.astype(np.float64)
    tree = cKDTree(P)
    d, j = tree.query(P, k=max_nn, distance_upper_bound=radius * (1 - 1e-12))
    checked = 0
    for i in range(len(P)):
        nb = j[i][np.isfinite(d[i])]
        if len(nb) < 3:
            assert out[i].tolist() == [0, 0, 1]
            continue
        q = P[nb]
        m = q.mean(0)
        cov = (q[:, :, None] * q[:, None, :]).mean(0) - np.outer(m, m)
        w, V = np.linalg.eigh(cov)
        if w[1] - w[0] < 1e-3 * w[2]:
            continue                                    # near-degenerate: the direction is not defined
        assert abs(abs(out[i].astype(np.float64) @ V[:, 0]) - 1.0) < 1e-5
        checked += 1
    assert checked > 0.6 * len(P)


def test_ransac_plane(oracle):
    r = np.random.default_rng(12)
    n = 5000
    pts = np.stack([r.uniform(-2, 2, n), 1.2 + r.normal(0, 0.003, n), r.uniform(0, 4, n)], axis=1)
    pts[:1500] = r.uniform(-2, 2, (1500, 3))
    pts = pts.astype(np.float32)
    plane, mask, best, counts = oracle.ransac_plane(pts, 0.01, 3, 200, seed=1234)
    assert abs(abs(plane[1]) - 1) < 1e-3 and abs(abs(plane[3]) - 1.2) < 5e-3
    assert best >= 0 and counts[best] == mask.sum() == counts.max()
    # inlier set = strict threshold on the winning hypothesis, restated independently
    ids = oracle.ransac_sample(1234, best, n, 3)
    p0, p1, p2 = pts[ids].astype(np.float64)
    nn = np.cross(p1 - p0, p2 - p0); nn /= np.linalg.norm(nn)
    dist = np.abs(pts.astype(np.float64) @ nn - nn @ p0)
    agree = (dist < 0.01) == mask.astype(bool)
    assert agree.mean() > 0.9999          # independent op order: only knife-edge points may differ
    # samples are distinct and reproducible
    assert len(set(ids.tolist())) == 3 and np.array_equal(ids, oracle.ransac_sample(1234, best, n, 3))
    # covariance fit for ransac_n > 3 (the reference's n = 30)
    # (with n = 30 a hypothesis is clean only if all 30 draws are inliers: use a band that is mostly floor)
    # NOTE upstream's early-exit bound log(1-p)/log(1-fitness^n) collapses to -inf once fitness^n < 1e-16
    # (n = 30: fitness < 0.29), which ends the search at the first such hypothesis; restated as is.
    band = pts[1500:]
    plane30, mask30, _, _ = oracle.ransac_plane(band, 0.01, 30, 100, seed=1234)
    assert abs(abs(plane30[1]) - 1) < 1e-2 and mask30.sum() > 0.95 * len(band)
    _, _, best_dirty, cnt_dirty = oracle.ransac_plane(pts, 0.01, 30, 100, seed=1234)
    assert best_dirty == 0 and cnt_dirty[0] / len(pts) < 0.29
    with pytest.raises(ValueError):
        oracle.ransac_plane(pts[:2], 0.01, 3, 10)


def test_ransac_plane_vs_independent_numpy_replay(oracle):
    """segment_plane written out independently in NumPy over the SAME counter-based samples: per-hypothesis plane
    (normalised cross product), strict-threshold inlier counts and rmse, upstream's sequential best rule with its
    early-exit bound, final inliers from the winning hypothesis, least-squares refit over them."""
    r = np.random.default_rng(12)
    n, H, thr, seed = 4000, 300, 0.01, 77
    pts = np.stack([r.uniform(-2, 2, n), 1.2 + r.normal(0, 0.003, n), r.uniform(0, 4, n)], axis=1)
    pts[:1500] = r.uniform(-2, 2, (1500, 3))
    pts = pts.astype(np.float32)
    plane, mask, best, counts = oracle.ransac_plane(pts, thr, 3, H, seed=seed)
    P = pts.astype(np.float64)
    cnt, rmse, planes = np.zeros(H, np.int64), np.zeros(H), np.zeros((H, 4))
    for h in range(H):
        p0, p1, p2 = P[oracle.ransac_sample(seed, h, n, 3)]
        nn = np.cross(p1 - p0, p2 - p0)
        if np.linalg.norm(nn) == 0:
            continue
        nn /= np.linalg.norm(nn)
        planes[h] = [*nn, -nn @ p0]
        e = np.abs(P @ nn - nn @ p0)
        cnt[h] = (e < thr).sum()
        rmse[h] = np.sqrt((e[e < thr] ** 2).sum() / max(cnt[h], 1))
    assert np.array_equal(cnt, counts)
    bf, br, bi, brk = 0.0, 0.0, -1, float(H)
    for h in range(H):
        if h >= brk:
            continue
        f = cnt[h] / n
        if cnt[h] and (f > bf or (f == bf and rmse[h] < br)):
            bf, br, bi = f, rmse[h], h
            brk = min(np.log(1 - 0.99999999) / np.log(1 - f ** 3), H) if f < 1 else 0
    assert bi == best and brk < H                        # the early exit was active in this case
    inl = np.abs(P @ planes[bi, :3] + planes[bi, 3]) < thr
    assert np.array_equal(inl, mask.astype(bool))
    q = P[inl]
    c = q.mean(0)
    _, _, Vt = np.linalg.svd(q - c)
    refit = np.array([*Vt[2], -Vt[2] @ c])
    if refit[:3] @ plane[:3] < 0:
        refit = -refit
    assert np.abs(refit - plane).max() < 1e-5            # upstream's determinant-based fit vs total least squares


def test_icp_recovers_planted_transform(oracle):
    from kinectpy_b200 import synth
    tgt = make_surface_cloud(6000, seed=13, outliers=0.0)
    nrm = oracle.estimate_normals(tgt, 0.1, 30)
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=1.0, shift_mm=(5, -5, 5), unit_scale=1e-3)
    src = oracle.transform(tgt[::2], np.linalg.inv(D))
    res = oracle.icp_point_to_plane(src, tgt, nrm, 0.05, init=np.eye(4), max_iter=30)
    assert res["fitness"] > 0.99
    assert np.abs(res["T"] - D).max() < 2e-3
    assert 1 <= res["iters"] <= 30
    # zero iterations: returns init and the initial correspondences' statistics
    r0 = oracle.icp_point_to_plane(src, tgt, nrm, 0.05, init=D, max_iter=0)
    assert np.array_equal(r0["T"], D) and r0["iters"] == 0


def _numpy_p2plane_icp(src, tgt, nrm, max_corr, init, max_iter, rel_fit=1e-6, rel_rmse=1e-6):
    """Open3D's RegistrationICP + TransformationEstimationPointToPlane written out independently:
    cKDTree 1-NN inside max_corr, r = (s - t).n, J = [s x n, n], JtJ x = -Jtr, T <- (Rz Ry Rx | t)(x) T, stop when both
    fitness and rmse move by less than the relative thresholds."""
    from scipy.spatial import cKDTree
    tree = cKDTree(tgt.astype(np.float64))
    T = np.array(init, np.float64)
    cur = src.astype(np.float64) @ T[:3, :3].T + T[:3, 3]

    def correspondences(p):
        d, j = tree.query(p, k=1, distance_upper_bound=max_corr)
        ok = np.isfinite(d) & (d < max_corr)
        return ok, j, d

    ok, j, d = correspondences(cur)
    fit, rmse = ok.mean(), (np.sqrt((d[ok] ** 2).mean()) if ok.any() else 0.0)
    iters = 0
    for it in range(max_iter):
        s, t, n = cur[ok], tgt[j[ok]].astype(np.float64), nrm[j[ok]].astype(np.float64)
        r = ((s - t) * n).sum(1)
        J = np.hstack([np.cross(s, n), n])
        x = np.linalg.solve(J.T @ J, -(J.T @ r))
        ca, sa, cb, sb, cg, sg = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
        U = np.eye(4)
        U[:3, :3] = [[cg * cb, cg * sb * sa - sg * ca, cg * sb * ca + sg * sa],
                     [sg * cb, sg * sb * sa + cg * ca, sg * sb * ca - cg * sa],
                     [-sb, cb * sa, cb * ca]]
        U[:3, 3] = x[3:]
        T = U @ T
        cur = cur @ U[:3, :3].T + U[:3, 3]
        ok, j, d = correspondences(cur)
        nfit, nrmse = ok.mean(), (np.sqrt((d[ok] ** 2).mean()) if ok.any() else 0.0)
        iters = it + 1
        done = abs(fit - nfit) < rel_fit and abs(rmse - nrmse) < rel_rmse
        fit, rmse = nfit, nrmse
        if done:
            break
    return T, fit, rmse, iters


def test_icp_vs_independent_numpy_restatement(oracle):
    """The oracle's ICP against a from-scratch NumPy / cKDTree transcription of the upstream loop: same iteration count,
    same fitness, transforms equal to rounding (the two differ only in summation order and in the linear solver)."""
    from kinectpy_b200 import synth
    tgt = make_surface_cloud(5000, seed=21, outliers=0.0)
    nrm = oracle.estimate_normals(tgt, 0.1, 30)
    D = synth.perturbed_extrinsic(np.eye(4), angle_deg=0.8, shift_mm=(4, -3, 5), unit_scale=1e-3)
    src = oracle.transform(tgt[1::2], np.linalg.inv(D))
    for max_corr, max_iter in ((0.05, 30), (0.02, 5)):
        res = oracle.icp_point_to_plane(src, tgt, nrm, max_corr, init=np.eye(4), max_iter=max_iter)
        T, fit, rmse, iters = _numpy_p2plane_icp(src, tgt, nrm, max_corr, np.eye(4), max_iter)
        assert res["iters"] == iters
        assert abs(res["fitness"] - fit) < 1e-12 and abs(res["rmse"] - rmse) < 1e-9
        assert np.abs(res["T"] - T).max() < 1e-9


def test_unproject_modes(oracle):
    from kinectpy_b200 import synth
    depth, tab, T = synth.render_sequence(synth.NFOV, 1, 2)
    xyz, valid, xyz16 = oracle.unproject(depth, tab, T, flags=oracle.F_INT16 | oracle.F_DROP_ANY_ZERO, scale=1.0,
                                         want_xyz16=True)
    P = synth.NFOV.pixels
    z = depth[0, 0].astype(np.float32)
    x = np.floor(tab[0, :, 0] * z + np.float32(0.5))
    ok = (~np.isnan(tab[0, :, 0])) & (depth[0, 0] > 0)
    assert np.array_equal(xyz16[0, :P, 0][ok], x[ok].astype(np.int16))
    assert np.array_equal(xyz16[0, :P, 2][ok], depth[0, 0][ok].astype(np.int16))
    assert np.all(xyz16[0, :P][~ok] == 0)
    # the reference's validity rule on the int16 triplets (utils/io.py:36)
    ref_valid = (xyz16[0] != 0).all(axis=1)
    assert np.array_equal(valid[0].astype(bool), ref_valid)
    assert np.isnan(xyz[0][~ref_valid]).all() and not np.isnan(xyz[0][ref_valid]).any()
    # sensor 1 rows are the transformed int16 points
    sub = xyz16[0, P:].astype(np.float64)
    exp = sub @ T[1, :3, :3].T + T[1, :3, 3]
    m = ref_valid[P:]
    assert np.allclose(xyz[0, P:][m], exp[m], atol=1e-3)
    # float mode, metres
    xyzf, validf, _ = oracle.unproject(depth, tab, None, flags=0, scale=1e-3)
    assert np.allclose(xyzf[0, :P, 2][ok], depth[0, 0][ok] * 1e-3, rtol=1e-7)
