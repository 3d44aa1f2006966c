"""Point-cloud container with the slice of the ``open3d.geometry.PointCloud`` interface that the
reference touches, backed by device buffers and the CUDA kernels behind
``libkinectpy_b200.so``.

Storage contract: points / colours / normals live on the GPU as float32 ``[n,3]``; ``.points``
materialises a float64 host copy (what ``np.asarray(pcd.points)`` gives with Open3D).  Every
geometric decision is taken in double on the float32-stored values.  There is no
CPU implementation: operations raise ``KinectPyB200Error`` without a CUDA device.
"""
from __future__ import annotations

import copy as _copy
import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _cabi
from ._cabi import DeviceArray, KinectPyB200Error

_SEED = [1234]   # the only seed the reference uses (train.py:19); see utility.random.seed


def _as_n3(a, name) -> np.ndarray:
    arr = np.asarray(a, dtype=np.float64)
    if arr.size == 0:
        return np.zeros((0, 3), dtype=np.float64)
    if arr.ndim != 2 or arr.shape[1] != 3:
        raise KinectPyB200Error(_cabi.KP_E_ARG, f"{name} must have shape (n, 3), got {arr.shape}")
    return np.ascontiguousarray(arr)


class _Attr:
    """One [n,3] attribute kept on the host (float64), on the device (float32), or both."""

    __slots__ = ("host", "dev", "n")

    def __init__(self):
        self.host: Optional[np.ndarray] = None
        self.dev: Optional[DeviceArray] = None
        self.n = 0

    def set_host(self, arr: np.ndarray):
        self.host, self.dev, self.n = arr, None, int(arr.shape[0])

    def set_dev(self, dev: Optional[DeviceArray], n: int):
        self.host, self.dev, self.n = None, dev, int(n)

    def present(self) -> bool:
        return self.n > 0 and (self.host is not None or self.dev is not None)

    def device(self, ctx, points: bool = False) -> Optional[DeviceArray]:
        if not self.present():
            return None
        if self.dev is None:
            h = self.host
            if points:
                # device convention: a row whose x is NaN is absent.  A row with ANY non-finite component (NaN in y or z,
                # an infinity) is uploaded as an all-NaN row, so every kernel skips it and remove_non_finite_points drops
                # it -- what Open3D's RemoveNonFinitePoints does; the host array is not modified
                bad = ~np.isfinite(h).all(axis=1)
                if bad.any():
                    h = h.copy()
                    h[bad] = np.nan
            self.dev = ctx.to_device(h, dtype=np.float32)
        return self.dev

    def peek_host(self) -> np.ndarray:
        """Host copy for READING: the device copy stays valid (internal callers that do not edit the array)."""
        if self.host is not None:
            return self.host
        if self.dev is None or self.n == 0:
            return np.zeros((0, 3), dtype=np.float64)
        return self.dev.to_host(self.n).astype(np.float64)

    def to_host(self) -> np.ndarray:
        if self.host is None:
            if self.dev is None or self.n == 0:
                self.host = np.zeros((0, 3), dtype=np.float64)
            else:
                self.host = self.dev.to_host(self.n).astype(np.float64)
            # the caller may now edit the array in place (Open3D semantics): the host copy rules
            self.dev = None
        return self.host


class PointCloud:
    """Duck-typed stand-in for ``o3d.geometry.PointCloud`` on the hot path."""

    def __init__(self, points=None, device: int = 0):
        self._device = device
        self._p, self._c, self._n = _Attr(), _Attr(), _Attr()
        if points is not None:
            self.points = points

    # ------------------------------------------------------------ plumbing
    @property
    def _ctx(self):
        return _cabi.default_context(self._device)

    @classmethod
    def _from_device(cls, ctx, pts: Optional[DeviceArray], n: int, colors=None, normals=None) -> "PointCloud":
        pc = cls(device=ctx.device)
        pc._p.set_dev(pts, n)
        if colors is not None:
            pc._c.set_dev(colors, n)
        if normals is not None:
            pc._n.set_dev(normals, n)
        return pc

    def _dev(self):
        ctx = self._ctx
        return ctx, self._p.device(ctx, points=True), self._c.device(ctx), self._n.device(ctx)

    @staticmethod
    def _ptr(d: Optional[DeviceArray]):
        return d.ptr if d is not None else None

    # ------------------------------------------------------- Open3D surface
    @property
    def points(self) -> np.ndarray:
        return self._p.to_host()

    @points.setter
    def points(self, v):
        self._p.set_host(_as_n3(v, "points"))

    @property
    def colors(self) -> np.ndarray:
        return self._c.to_host()

    @colors.setter
    def colors(self, v):
        self._c.set_host(_as_n3(v, "colors"))

    @property
    def normals(self) -> np.ndarray:
        return self._n.to_host()

    @normals.setter
    def normals(self, v):
        self._n.set_host(_as_n3(v, "normals"))

    def __len__(self):
        return self._p.n

    def has_points(self) -> bool:
        return self._p.n > 0

    def has_colors(self) -> bool:
        return self._p.n > 0 and self._c.n == self._p.n

    def has_normals(self) -> bool:
        return self._p.n > 0 and self._n.n == self._p.n

    def is_empty(self) -> bool:
        return self._p.n == 0

    def __repr__(self):
        return f"PointCloud with {self._p.n} points."

    def __deepcopy__(self, memo):
        out = PointCloud(device=self._device)
        for src, dst in ((self._p, out._p), (self._c, out._c), (self._n, out._n)):
            if src.host is not None:
                dst.set_host(src.host.copy())
            elif src.dev is not None:
                dst.set_dev(src.dev.copy(src.n), src.n)
        return out

    def clone(self) -> "PointCloud":
        return _copy.deepcopy(self)

    def get_min_bound(self) -> np.ndarray:
        return self._bounds()[:3].astype(np.float64)

    def get_max_bound(self) -> np.ndarray:
        return self._bounds()[3:].astype(np.float64)

    def _bounds(self) -> np.ndarray:
        if self._p.n == 0:
            return np.zeros(6, dtype=np.float32)
        ctx, pts, _, _ = self._dev()
        b = (C.c_float * 6)()
        nv = C.c_int64()
        ctx.check(ctx.lib.kp_bounds(ctx.handle, pts.ptr, self._p.n, b, C.byref(nv)))
        return np.array(list(b), dtype=np.float32)

    def paint_uniform_color(self, color):
        col = np.asarray(color, dtype=np.float64).reshape(3)
        self._c.set_host(np.tile(col, (self._p.n, 1)))
        return self

    def transform(self, T) -> "PointCloud":
        """In-place ``p' = R p + t`` (reference call: preprocessing/data.py:46-48); normals rotate."""
        if self._p.n == 0:
            return self
        ctx, pts, _, nrm = self._dev()
        t16 = _cabi.T16(T)
        ctx.check(ctx.lib.kp_transform_points(ctx.handle, pts.ptr, self._p.n, t16.ctypes.data, 0))
        if nrm is not None and self.has_normals():
            ctx.check(ctx.lib.kp_transform_points(ctx.handle, nrm.ptr, self._n.n, t16.ctypes.data, 1))
            self._n.host = None
        self._p.host = None
        return self

    def __add__(self, other: "PointCloud") -> "PointCloud":
        """Concatenation, left operand first (floor_removal.py:72); attributes kept only if both have them."""
        ctx = self._ctx
        n1, n2 = self._p.n, other._p.n
        n = n1 + n2
        out = PointCloud(device=self._device)
        if n == 0:
            return out

        def cat(a: _Attr, b: _Attr):
            d = ctx.empty((n, 3), np.float32)
            if n1:
                ctx.check(ctx.lib.kp_memcpy_d2d(ctx.handle, d.ptr, a.device(ctx).ptr, n1 * 12))
            if n2:
                ctx.check(ctx.lib.kp_memcpy_d2d(ctx.handle, d.ptr + n1 * 12, b.device(ctx).ptr, n2 * 12))
            return d

        out._p.set_dev(cat(self._p, other._p), n)
        both = lambda a, b, na, nb: (na == 0 or a.n == na) and (nb == 0 or b.n == nb) and (a.n + b.n == n)
        if both(self._c, other._c, n1, n2):
            out._c.set_dev(cat(self._c, other._c), n)
        if both(self._n, other._n, n1, n2):
            out._n.set_dev(cat(self._n, other._n), n)
        return out

    # ---- voxel ---------------------------------------------------------
    def voxel_down_sample(self, voxel_size: float) -> "PointCloud":
        """``PointCloud.voxel_down_sample`` (preprocessing/filtering.py:23, registration.py:8)."""
        out, _, _ = self._voxel(voxel_size, want_maps=False)
        return out

    def voxel_down_sample_and_trace(self, voxel_size: float):
        """Returns (cloud, ijk int32[m,3], point_voxel int32[n]) -- the bit-exact artefacts."""
        return self._voxel(voxel_size, want_maps=True)

    def _voxel(self, voxel_size, want_maps):
        if not voxel_size > 0:
            raise KinectPyB200Error(_cabi.KP_E_ARG, "voxel_size <= 0")
        n = self._p.n
        if n == 0:
            return PointCloud(device=self._device), np.zeros((0, 3), np.int32), np.zeros((0,), np.int32)
        ctx, pts, col, nrm = self._dev()
        col = col if self.has_colors() else None
        nrm = nrm if self.has_normals() else None
        o_p = ctx.empty((n, 3), np.float32)
        o_c = ctx.empty((n, 3), np.float32) if col is not None else None
        o_n = ctx.empty((n, 3), np.float32) if nrm is not None else None
        ijk = ctx.empty((n, 3), np.int32) if want_maps else None
        pv = ctx.empty((n,), np.int32) if want_maps else None
        m = C.c_int64()
        ctx.check(ctx.lib.kp_voxel_downsample(ctx.handle, pts.ptr, self._ptr(col), self._ptr(nrm), n, float(voxel_size),
                                              o_p.ptr, self._ptr(o_c), self._ptr(o_n), self._ptr(ijk), self._ptr(pv),
                                              None, C.byref(m)))
        out = PointCloud._from_device(ctx, o_p, m.value, o_c, o_n)
        if m.value <= 0.8 * n:   # the voxel grid, not the sensor, now sets the point spacing: remember it
            out._voxel_hint = float(voxel_size)
        if want_maps:
            return out, ijk.to_host(m.value), pv.to_host()
        return out, None, None

    # ---- masks / selection ----------------------------------------------
    def _select_mask(self, mask_dev: DeviceArray, invert=False, want_index=True):
        ctx, pts, col, nrm = self._dev()
        n = self._p.n
        col = col if self.has_colors() else None
        nrm = nrm if self.has_normals() else None
        o_p = ctx.empty((n, 3), np.float32)
        o_c = ctx.empty((n, 3), np.float32) if col is not None else None
        o_n = ctx.empty((n, 3), np.float32) if nrm is not None else None
        idx = ctx.empty((n,), np.int32) if want_index else None
        cnt = C.c_int64()
        ctx.check(ctx.lib.kp_compact(ctx.handle, n, mask_dev.ptr, 1 if invert else 0, pts.ptr, o_p.ptr, self._ptr(col),
                                     self._ptr(o_c), self._ptr(nrm), self._ptr(o_n), self._ptr(idx), C.byref(cnt)))
        out = PointCloud._from_device(ctx, o_p, cnt.value, o_c, o_n)
        if hasattr(self, "_voxel_hint"):
            out._voxel_hint = self._voxel_hint
        return out, (idx.to_host(cnt.value) if want_index else None)

    def select_by_index(self, indices, invert: bool = False) -> "PointCloud":
        """``select_by_index`` (floor_removal.py:50,69,71,72).  The reference passes ascending, unique index lists
        (the results of ``remove_statistical_outlier`` / ``segment_plane``); those go through a mask and an ordered
        compaction on the device.  Any other list (unsorted, repeated indices) is gathered in the order given, as
        Open3D does; ``invert`` ignores order and repetition by definition."""
        n = self._p.n
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        if idx.size and (idx.min() < 0 or idx.max() >= n):
            raise KinectPyB200Error(_cabi.KP_E_ARG, "select_by_index: index out of range")
        if not invert and idx.size > 1 and not bool(np.all(idx[1:] > idx[:-1])):
            out = PointCloud(device=self._device)
            out._p.set_host(self._p.peek_host()[idx].copy())
            if self.has_colors():
                out._c.set_host(self._c.peek_host()[idx].copy())
            if self.has_normals():
                out._n.set_host(self._n.peek_host()[idx].copy())
            return out
        mask = np.zeros(n, dtype=np.uint8)
        mask[idx] = 1
        if n == 0:
            return PointCloud(device=self._device)
        out, _ = self._select_mask(self._ctx.to_device(mask), invert=invert, want_index=False)
        return out

    def remove_non_finite_points(self) -> "PointCloud":
        n = self._p.n
        if n == 0:
            return self
        ctx, pts, col, nrm = self._dev()
        o_p = ctx.empty((n, 3), np.float32)
        cnt = C.c_int64()
        col = col if self.has_colors() else None
        o_c = ctx.empty((n, 3), np.float32) if col is not None else None
        ctx.check(ctx.lib.kp_compact(ctx.handle, n, None, 0, pts.ptr, o_p.ptr, self._ptr(col), self._ptr(o_c), None, None,
                                     None, C.byref(cnt)))
        self._p.set_dev(o_p, cnt.value)
        if o_c is not None:
            self._c.set_dev(o_c, cnt.value)
        self._n = _Attr()
        return self

    def remove_statistical_outlier(self, nb_neighbors: int, std_ratio: float, print_progress: bool = False):
        """``remove_statistical_outlier`` (preprocessing/filtering.py:24, floor_removal.py:73)."""
        if nb_neighbors < 1 or not std_ratio > 0:
            raise KinectPyB200Error(_cabi.KP_E_ARG, "nb_neighbors < 1 or std_ratio <= 0")
        n = self._p.n
        if n == 0:
            return PointCloud(device=self._device), []
        ctx, pts, _, _ = self._dev()
        keep = ctx.empty((n,), np.uint8)
        kept = C.c_int64()
        stats = (C.c_double * 3)()
        hint = 0.0
        if hasattr(self, "_voxel_hint"):   # cloud came out of voxel_down_sample: its spacing sizes the search grid
            hint = self._voxel_hint * 1.5 * float(np.sqrt(nb_neighbors / np.pi))
        ctx.check(ctx.lib.kp_sor_mask(ctx.handle, pts.ptr, n, int(nb_neighbors), float(std_ratio), hint, keep.ptr, None,
                                      stats, C.byref(kept)))
        out, idx = self._select_mask(keep)
        self._last_sor_stats = tuple(stats)
        return out, idx.tolist()

    def remove_radius_outlier(self, nb_points: int, radius: float, print_progress: bool = False):
        """``remove_radius_outlier`` (Open3D semantics)."""
        if nb_points < 1 or not radius > 0:
            raise KinectPyB200Error(_cabi.KP_E_ARG, "nb_points < 1 or radius <= 0")
        n = self._p.n
        if n == 0:
            return PointCloud(device=self._device), []
        ctx, pts, _, _ = self._dev()
        keep = ctx.empty((n,), np.uint8)
        kept = C.c_int64()
        ctx.check(ctx.lib.kp_radius_mask(ctx.handle, pts.ptr, n, int(nb_points), float(radius), keep.ptr, None, C.byref(kept)))
        out, idx = self._select_mask(keep)
        return out, idx.tolist()

    # ---- normals ---------------------------------------------------------
    def estimate_normals(self, search_param=None, fast_normal_computation: bool = True):
        """``estimate_normals(KDTreeSearchParamHybrid(radius, max_nn))`` (registration.py:11-13)."""
        sp = search_param if search_param is not None else KDTreeSearchParamKNN(30)
        radius = float(getattr(sp, "radius", 0.0) or 0.0)
        max_nn = int(getattr(sp, "max_nn", getattr(sp, "knn", 30)))
        n = self._p.n
        if n == 0:
            return self
        ctx, pts, _, _ = self._dev()
        nrm = ctx.empty((n, 3), np.float32)
        ctx.check(ctx.lib.kp_estimate_normals(ctx.handle, pts.ptr, n, radius, max_nn, nrm.ptr))
        self._n.set_dev(nrm, n)
        return self

    # ---- RANSAC ----------------------------------------------------------
    def segment_plane(self, distance_threshold: float, ransac_n: int, num_iterations: int,
                      probability: float = 0.99999999, seed: Optional[int] = None):
        """``segment_plane`` (floor_removal.py:70) -> (plane float64[4], inlier index list)."""
        n = self._p.n
        ctx, pts, _, _ = self._dev()
        if pts is None:
            raise KinectPyB200Error(_cabi.KP_E_ARG, "segment_plane on an empty cloud")
        mask = ctx.empty((n,), np.uint8)
        plane = (C.c_double * 4)()
        ninl = C.c_int64()
        best = C.c_int32()
        ctx.check(ctx.lib.kp_ransac_plane(ctx.handle, pts.ptr, n, float(distance_threshold), int(ransac_n),
                                          int(num_iterations), float(probability),
                                          int(_SEED[0] if seed is None else seed) & 0xFFFFFFFFFFFFFFFF, plane, mask.ptr,
                                          C.byref(ninl), C.byref(best), None))
        inliers = np.flatnonzero(mask.to_host()).tolist()
        return np.array(list(plane), dtype=np.float64), inliers


# ----------------------------------------------------------- search params --
class KDTreeSearchParamHybrid:
    def __init__(self, radius: float, max_nn: int):
        self.radius, self.max_nn = float(radius), int(max_nn)


class KDTreeSearchParamKNN:
    def __init__(self, knn: int = 30):
        self.knn = int(knn)
        self.max_nn = int(knn)
        self.radius = 0.0


def Vector3dVector(a) -> np.ndarray:
    """``o3d.utility.Vector3dVector``: the container here is simply a float64 ``(n,3)`` array."""
    return _as_n3(a, "Vector3dVector")


def seed(value: int):
    """``o3d.utility.random.seed``: seeds the RANSAC hypothesis generator."""
    _SEED[0] = int(value)


# ------------------------------------------------------------ registration --
class TransformationEstimationPointToPlane:
    pass


class TransformationEstimationPointToPoint:
    def __init__(self, with_scaling: bool = False):
        if with_scaling:
            raise NotImplementedError("with_scaling=True is not used by the reference")
        self.with_scaling = with_scaling

    def compute_transformation(self, source: "PointCloud", target: "PointCloud", corres) -> np.ndarray:
        """``p2p.compute_transformation(source, target, Vector2iVector(corr))`` at
        ``manual_pointcloud_registration.py:90-92``: Umeyama / Kabsch over a handful of hand-picked pairs
        (host-side; three to a dozen points are not device work)."""
        c = np.asarray(corres).astype(np.int64).reshape(-1, 2)
        # read-only view of our own clouds (their device copies stay valid); anything else with a `.points` array works too
        rows = lambda pc: pc._p.peek_host() if isinstance(pc, PointCloud) else np.asarray(pc.points)
        s = rows(source)[c[:, 0]]
        t = rows(target)[c[:, 1]]
        T = np.eye(4)
        if len(c) == 0:
            return T
        ms, mt = s.mean(axis=0), t.mean(axis=0)
        cov = (t - mt).T @ (s - ms) / len(c)
        U, _, Vt = np.linalg.svd(cov)
        S = np.diag([1.0, 1.0, -1.0 if np.linalg.det(U) * np.linalg.det(Vt) < 0 else 1.0])
        R = U @ S @ Vt
        T[:3, :3] = R
        T[:3, 3] = mt - R @ ms
        return T


class TransformationEstimationForColoredICP:
    def __init__(self, lambda_geometric: float = 0.968):
        self.lambda_geometric = float(lambda_geometric)


def Vector2iVector(a) -> np.ndarray:
    return np.asarray(a).astype(np.int32).reshape(-1, 2)


class ICPConvergenceCriteria:
    def __init__(self, relative_fitness: float = 1e-6, relative_rmse: float = 1e-6, max_iteration: int = 30):
        self.relative_fitness, self.relative_rmse, self.max_iteration = relative_fitness, relative_rmse, max_iteration


class RegistrationResult:
    def __init__(self, T, fitness, rmse, iters, ncorr, correspondence_set=None):
        self.correspondence_set = correspondence_set
        self.transformation = T
        self.fitness = fitness
        self.inlier_rmse = rmse
        self.num_iterations = iters
        self.num_correspondences = ncorr

    def __repr__(self):
        return (f"RegistrationResult with fitness={self.fitness:e}, inlier_rmse={self.inlier_rmse:e}, "
                f"and correspondence_set size of {self.num_correspondences}")


def _icp_outputs():
    return np.zeros(16, dtype=np.float64), C.c_double(), C.c_double(), C.c_int(), C.c_int64()


def registration_icp(source: PointCloud, target: PointCloud, max_correspondence_distance: float, init=None,
                     estimation_method=None, criteria: Optional[ICPConvergenceCriteria] = None) -> RegistrationResult:
    """``o3d.pipelines.registration.registration_icp``: point-to-plane (``registration.py:78-84``) and
    point-to-point (``manual_pointcloud_registration.py:94-98``; Open3D's default estimator)."""
    if estimation_method is None:
        estimation_method = TransformationEstimationPointToPoint()
    if isinstance(estimation_method, TransformationEstimationForColoredICP):
        return registration_colored_icp(source, target, max_correspondence_distance, init, estimation_method, criteria)
    if not max_correspondence_distance > 0:
        raise KinectPyB200Error(_cabi.KP_E_ARG, "max_correspondence_distance <= 0")
    crit = criteria or ICPConvergenceCriteria()
    ctx = source._ctx
    _, s_pts, _, _ = source._dev()
    _, t_pts, _, t_nrm = target._dev()
    T0 = _cabi.T16(np.eye(4) if init is None else init)
    T, fit, rmse, iters, nc = _icp_outputs()
    if isinstance(estimation_method, TransformationEstimationPointToPlane):
        if not target.has_normals():
            raise KinectPyB200Error(_cabi.KP_E_ARG, "TransformationEstimationPointToPlane requires target normals")
        ctx.check(ctx.lib.kp_icp_point_to_plane(ctx.handle, PointCloud._ptr(s_pts), len(source), PointCloud._ptr(t_pts),
                                                PointCloud._ptr(t_nrm), len(target), float(max_correspondence_distance),
                                                T0.ctypes.data, int(crit.max_iteration), float(crit.relative_fitness),
                                                float(crit.relative_rmse), T.ctypes.data, C.byref(fit), C.byref(rmse),
                                                C.byref(iters), C.byref(nc)))
    elif isinstance(estimation_method, TransformationEstimationPointToPoint):
        ctx.check(ctx.lib.kp_icp_point_to_point(ctx.handle, PointCloud._ptr(s_pts), len(source), PointCloud._ptr(t_pts),
                                                len(target), float(max_correspondence_distance), T0.ctypes.data,
                                                int(crit.max_iteration), float(crit.relative_fitness),
                                                float(crit.relative_rmse), T.ctypes.data, C.byref(fit), C.byref(rmse),
                                                C.byref(iters), C.byref(nc)))
    else:
        raise TypeError("unknown estimation method %r" % (estimation_method,))
    return RegistrationResult(T.reshape(4, 4).copy(), fit.value, rmse.value, iters.value, nc.value)


def registration_colored_icp(source: PointCloud, target: PointCloud, max_correspondence_distance: float, init=None,
                             estimation_method: Optional[TransformationEstimationForColoredICP] = None,
                             criteria: Optional[ICPConvergenceCriteria] = None) -> RegistrationResult:
    """``o3d.pipelines.registration.registration_colored_icp`` (``registration.py:108-113``)."""
    est = estimation_method or TransformationEstimationForColoredICP()
    if not max_correspondence_distance > 0:
        raise KinectPyB200Error(_cabi.KP_E_ARG, "max_correspondence_distance <= 0")
    if not (source.has_colors() and target.has_colors() and target.has_normals()):
        raise KinectPyB200Error(_cabi.KP_E_ARG, "ColoredICP requires source colours and target colours + normals")
    crit = criteria or ICPConvergenceCriteria()
    ctx = source._ctx
    _, s_pts, s_col, _ = source._dev()
    _, t_pts, t_col, t_nrm = target._dev()
    T0 = _cabi.T16(np.eye(4) if init is None else init)
    T, fit, rmse, iters, nc = _icp_outputs()
    ctx.check(ctx.lib.kp_icp_colored(ctx.handle, PointCloud._ptr(s_pts), PointCloud._ptr(s_col), len(source),
                                     PointCloud._ptr(t_pts), PointCloud._ptr(t_col), PointCloud._ptr(t_nrm), len(target),
                                     float(max_correspondence_distance), est.lambda_geometric, T0.ctypes.data,
                                     int(crit.max_iteration), float(crit.relative_fitness), float(crit.relative_rmse),
                                     T.ctypes.data, C.byref(fit), C.byref(rmse), C.byref(iters), C.byref(nc)))
    return RegistrationResult(T.reshape(4, 4).copy(), fit.value, rmse.value, iters.value, nc.value)


# ----------------------------------------------------- global registration --
class Feature:
    """``o3d.pipelines.registration.Feature``: ``data`` is ``float64 (dimension, num)`` like Open3D's; the
    device copy (``float64 [num][dimension]``) is kept for the matching kernels."""

    def __init__(self, ctx=None, dev: Optional[DeviceArray] = None, num: int = 0, dim: int = 33):
        self._ctx, self._devarr, self._num, self._dim, self._host = ctx, dev, int(num), int(dim), None

    def dimension(self) -> int:
        return self._dim

    def num(self) -> int:
        return self._num

    @property
    def data(self) -> np.ndarray:
        if self._host is None:
            self._host = (self._devarr.to_host(self._num).T.copy() if self._devarr is not None
                          else np.zeros((self._dim, 0), np.float64))
        return self._host

    def _device(self, ctx) -> DeviceArray:
        if self._devarr is None:
            self._devarr = ctx.to_device(np.ascontiguousarray(self.data.T), np.float64)
        return self._devarr


def compute_fpfh_feature(input: PointCloud, search_param) -> Feature:
    """``o3d.pipelines.registration.compute_fpfh_feature`` (``registration.py:17-20``)."""
    if not input.has_normals():
        raise KinectPyB200Error(_cabi.KP_E_ARG, "compute_fpfh_feature: the cloud has no normals")
    radius = float(getattr(search_param, "radius", 0.0) or 0.0)
    max_nn = int(getattr(search_param, "max_nn", getattr(search_param, "knn", 100)))
    ctx, pts, _, nrm = input._dev()
    n = len(input)
    feat = ctx.empty((max(n, 1), 33), np.float64)
    ctx.check(ctx.lib.kp_fpfh(ctx.handle, PointCloud._ptr(pts), PointCloud._ptr(nrm), n, radius, max_nn, feat.ptr))
    return Feature(ctx, feat, n, 33)


class CorrespondenceCheckerBasedOnEdgeLength:
    def __init__(self, similarity_threshold: float = 0.9):
        self.similarity_threshold = float(similarity_threshold)


class CorrespondenceCheckerBasedOnDistance:
    def __init__(self, distance_threshold: float):
        self.distance_threshold = float(distance_threshold)


class CorrespondenceCheckerBasedOnNormal:
    def __init__(self, normal_angle_threshold: float):
        raise NotImplementedError("CorrespondenceCheckerBasedOnNormal is not used by the reference")


class RANSACConvergenceCriteria:
    def __init__(self, max_iteration: int = 100000, confidence: float = 0.999):
        self.max_iteration, self.confidence = int(max_iteration), float(confidence)


_RANSAC_CALLS = [0]   # successive calls draw different hypothesis streams, as successive upstream calls advance the global RNG


def registration_ransac_based_on_feature_matching(source: PointCloud, target: PointCloud, source_feature: Feature,
                                                  target_feature: Feature, mutual_filter: bool,
                                                  max_correspondence_distance: float, estimation_method=None,
                                                  ransac_n: int = 3, checkers=(), criteria=None, seed: Optional[int] = None
                                                  ) -> RegistrationResult:
    """``o3d.pipelines.registration.registration_ransac_based_on_feature_matching`` (``registration.py:50-57``):
    nearest feature of every source point among the target's (and back, for the mutual filter), then RANSAC
    over those correspondences.  Matching, hypothesis checking and scoring run on the GPU; the mutual filter is
    an index comparison on the host."""
    if estimation_method is not None and not isinstance(estimation_method, TransformationEstimationPointToPoint):
        raise NotImplementedError("only TransformationEstimationPointToPoint(False) is used by the reference")
    crit = criteria or RANSACConvergenceCriteria()
    if ransac_n < 3 or not max_correspondence_distance > 0:
        return RegistrationResult(np.eye(4), 0.0, 0.0, 0, 0, np.zeros((0, 2), np.int32))
    edge, dist = 0.0, 0.0
    for ch in checkers:
        if isinstance(ch, CorrespondenceCheckerBasedOnEdgeLength):
            edge = ch.similarity_threshold
        elif isinstance(ch, CorrespondenceCheckerBasedOnDistance):
            dist = ch.distance_threshold
        else:
            raise NotImplementedError("unsupported correspondence checker %r" % (ch,))
    ctx, s_pts, _, _ = source._dev()
    _, t_pts, _, _ = target._dev()
    ns, nt = len(source), len(target)
    fs, ft = source_feature._device(ctx), target_feature._device(ctx)
    nn_st = ctx.empty((max(ns, 1),), np.int32)
    ctx.check(ctx.lib.kp_feature_match(ctx.handle, fs.ptr, ns, ft.ptr, nt, 33, nn_st.ptr, None))
    st = nn_st.to_host(ns)
    corres = np.stack([np.arange(ns, dtype=np.int32), st], axis=1)
    if mutual_filter:
        nn_ts = ctx.empty((max(nt, 1),), np.int32)
        ctx.check(ctx.lib.kp_feature_match(ctx.handle, ft.ptr, nt, fs.ptr, ns, 33, nn_ts.ptr, None))
        ts = nn_ts.to_host(nt)
        mut = corres[ts[st] == np.arange(ns)] if ns and nt else corres[:0]
        if len(mut) >= ransac_n * 3:
            corres = mut                                   # else: too few, fall back to the unfiltered set (upstream)
    corres = np.ascontiguousarray(corres, dtype=np.int32)
    if seed is None:
        seed = (_SEED[0] + 0x9E3779B97F4A7C15 * _RANSAC_CALLS[0]) & 0xFFFFFFFFFFFFFFFF
        _RANSAC_CALLS[0] += 1
    d_cor = ctx.to_device(corres, np.int32)
    T = np.zeros(16, np.float64)
    fit, rmse, best, val = C.c_double(), C.c_double(), C.c_int32(), C.c_int64()
    ctx.check(ctx.lib.kp_ransac_correspondence(ctx.handle, PointCloud._ptr(s_pts), ns, PointCloud._ptr(t_pts), nt, d_cor.ptr,
                                               len(corres), float(max_correspondence_distance), int(ransac_n), edge, dist,
                                               crit.max_iteration, crit.confidence, int(seed), T.ctypes.data_as(C.POINTER(C.c_double)),
                                               C.byref(fit), C.byref(rmse), C.byref(best), C.byref(val)))
    T = T.reshape(4, 4).copy()
    # correspondence_set of the result: the correspondences the winning transform brings within the threshold
    inl = corres[:0]
    if fit.value > 0 and len(corres):
        sp = source._p.peek_host()[corres[:, 0]] @ T[:3, :3].T + T[:3, 3]     # (read-only: no re-upload on the next trial)
        d2 = ((sp - target._p.peek_host()[corres[:, 1]]) ** 2).sum(axis=1)
        inl = corres[d2 < max_correspondence_distance ** 2]
    res = RegistrationResult(T, fit.value, rmse.value, int(best.value), len(inl), inl)
    res.num_validated = int(val.value)
    return res
