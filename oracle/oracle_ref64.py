"""O-ref: the float64-STORAGE tier of the oracle -- TEST INFRASTRUCTURE, not product.

The reference keeps every cloud as float64 (``utils/io.py:29``, Open3D's ``Vector3dVector``); the product (and
``oracle/kp_oracle.c``, "O-bits") stores float32 and decides in double on those stored values.  This module restates
the same decisions with float64 storage end to end -- fused coordinates, voxel means, neighbour distances, plane
distances -- in plain NumPy / SciPy, so that the RATE at which a decision differs between the two storage
precisions can be measured (``tests/test_oracle_tiers.py``).  Semantics follow SURVEY.md Appendix A:
A.1 depth -> XYZ (k4a rounding), A.8 transform, A.2 VoxelDownSample (``preprocessing/filtering.py:23``),
A.3 RemoveStatisticalOutliers (``filtering.py:24``), A.5 SegmentPlane (``floor_removal.py:70``; hypotheses from the
shared counter-based generator).  Only tests/ may import this file."""
from __future__ import annotations

import numpy as np


def fuse64(depth, xytab, T, drop_any_zero=True, scale=1e-3):
    """depth uint16[S,P], table float32[S,P,2], T float64[S,4,4] -> float64[S*P,3] (NaN = invalid), sensor-major."""
    S, P = depth.shape
    out = np.full((S, P, 3), np.nan, np.float64)
    for s in range(S):
        xt, yt = xytab[s, :, 0], xytab[s, :, 1]
        z = depth[s].astype(np.float32)
        ok = ~(np.isnan(xt) | np.isnan(yt)) & (depth[s] != 0)
        with np.errstate(invalid="ignore"):
            xi = np.floor(xt * z + np.float32(0.5)).astype(np.float32)        # fp32 product, round half up (k4a)
            yi = np.floor(yt * z + np.float32(0.5)).astype(np.float32)
        X = np.where(ok, xi, 0).astype(np.int16).astype(np.float64) * scale
        Y = np.where(ok, yi, 0).astype(np.int16).astype(np.float64) * scale
        Z = np.where(ok, depth[s], 0).astype(np.int16).astype(np.float64) * scale
        if drop_any_zero:
            ok &= (X != 0) & (Y != 0) & (Z != 0)
        p = np.stack([X, Y, Z], 1)
        R, t = np.asarray(T[s], np.float64)[:3, :3], np.asarray(T[s], np.float64)[:3, 3]
        q = np.empty_like(p)
        for r in range(3):        # ((a x + b y) + c z) + d, the order of the product's kp_affine
            q[:, r] = ((R[r, 0] * p[:, 0] + R[r, 1] * p[:, 1]) + R[r, 2] * p[:, 2]) + t[r]
        out[s][ok] = q[ok]
    return out.reshape(S * P, 3)


def voxel64(xyz, voxel):
    """float64 rows (NaN = absent) -> dict(points float64[M,3] in canonical (ix,iy,iz) order, ijk int64[M,3],
    point_ijk int64[N,3] (-1 rows for absent points), min_bound)."""
    ok = ~np.isnan(xyz[:, 0])
    p = xyz[ok]
    minb = p.min(0) - voxel * 0.5
    ijk = np.floor((p - minb) / voxel).astype(np.int64)
    order = np.lexsort((ijk[:, 2], ijk[:, 1], ijk[:, 0]))          # stable: points of a voxel stay in input order
    s = ijk[order]
    head = np.ones(len(s), bool)
    head[1:] = (s[1:] != s[:-1]).any(1)
    starts = np.flatnonzero(head)
    sums = np.add.reduceat(p[order], starts, axis=0)              # (pairwise inside numpy: tolerance-level only)
    cnt = np.diff(np.append(starts, len(s)))[:, None]
    point_ijk = np.full((len(xyz), 3), -1, np.int64)
    point_ijk[ok] = ijk
    return {"points": sums / cnt, "ijk": s[starts], "point_ijk": point_ijk, "min_bound": minb}


def sor64(points, k, std_ratio):
    from scipy.spatial import cKDTree
    n = len(points)
    kk = min(k, n)
    d, _ = cKDTree(points).query(points, k=kk, workers=-1)
    d = d.reshape(n, kk)
    mean = d.sum(1) / kk
    pos = mean > 0
    mu = mean[pos].sum() / n
    sd = np.sqrt(((mean[pos] - mu) ** 2).sum() / (n - 1))
    return (pos & (mean < mu + std_ratio * sd)), mean, (mu, sd)


def ransac64(points, thr, iters, samples, probability=0.99999999):
    """samples int64[iters,3] (the shared generator's index triples); returns (best hypothesis, inlier mask, counts)."""
    n = len(points)
    counts = np.zeros(iters, np.int64)
    planes = np.zeros((iters, 4))
    valid = np.zeros(iters, bool)
    for h in range(iters):
        p0, p1, p2 = points[samples[h]]
        nrm = np.cross(p1 - p0, p2 - p0)
        nn = np.sqrt((nrm * nrm).sum())
        if nn == 0 or np.isnan(nn):
            continue
        nrm = nrm / nn
        planes[h] = (*nrm, -(nrm * p0).sum())
        valid[h] = True
        counts[h] = int((np.abs(points @ nrm + planes[h, 3]) < thr).sum())
    best, best_fit, break_it, done = -1, 0.0, float(iters), 0
    for h in range(iters):
        if done > break_it:
            continue
        if not valid[h]:
            continue
        fit = counts[h] / n
        if fit > best_fit:
            best, best_fit = h, fit
            break_it = min(np.log(1 - probability) / np.log(1 - fit ** 3), iters) if fit < 1 else 0
        done += 1
    mask = np.abs(points @ planes[best, :3] + planes[best, 3]) < thr if best >= 0 else np.zeros(n, bool)
    return best, mask, counts
