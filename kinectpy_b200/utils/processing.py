"""B200 counterpart of the point-cloud helpers of the reference's ``utils/processing.py`` that sit on the
hot path: ``select_points_randomly`` (``utils/processing.py:259-275``, the fixed-N resampling that feeds
PointNet at ``datasets/kinect_dataset.py:103-104``), the ``points[:N]`` prefix variant of
``datasets/kinect_dataset_npz.py:96-97`` and ``statistical_outlier_removal``
(``utils/processing.py:302-310``).  Same names, argument order and defaults.

``np.random.choice`` draws from numpy's global state; here the subset is a pure function of
``(seed, stream, index)`` (see ``include/kp_api.h``, K6), so it is reproducible on any number of GPUs.
``seed(value)`` plays the role of ``np.random.seed``; every call without an explicit ``stream`` consumes
the next stream number, as successive ``np.random.choice`` calls consume the global generator.
"""
from __future__ import annotations

import ctypes as C
import itertools
from typing import Optional, Sequence

import numpy as np

from .. import _cabi
from ..geometry import PointCloud

_state = {"seed": 1234, "stream": itertools.count()}   # 1234: the only seed the reference sets (train.py:19)


def seed(value: int) -> None:
    _state["seed"] = int(value) & 0xFFFFFFFFFFFFFFFF
    _state["stream"] = itertools.count()


def select_points_randomly(pointcloud, number_of_points: int, stream: Optional[int] = None,
                           return_index: bool = False) -> np.ndarray:
    """``number_of_points`` distinct points of the cloud, uniformly at random, as ``float64 (N, 3)``.
    Raises ``ValueError`` when the cloud has fewer points (``np.random.choice(..., replace=False)`` does)."""
    pts = pointcloud.points if isinstance(pointcloud, PointCloud) else np.asarray(pointcloud)
    n = int(pts.shape[0])
    N = int(number_of_points)
    if N > n:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    if N < 0:
        raise ValueError("negative dimensions are not allowed")
    ctx = _cabi.default_context()
    d = ctx.to_device(pts, np.float32)
    out = ctx.empty((N, 3), np.float32)
    idx = ctx.empty((N,), np.int32)
    cnt = C.c_int64()
    s = next(_state["stream"]) if stream is None else int(stream)
    ctx.check(ctx.lib.kp_resample_fixed_n(ctx.handle, d.ptr, n, N, _cabi.RESAMPLE_RANDOM, _state["seed"], s, out.ptr, idx.ptr,
                                          C.byref(cnt)))
    sel = idx.to_host()
    res = np.asarray(pts, dtype=np.float64)[sel]          # the reference returns rows of the float64 points array
    return (res, sel) if return_index else res


def take_first_points(points, number_of_points: int) -> np.ndarray:
    """``npz_file['points'][:number_of_points]`` (``datasets/kinect_dataset_npz.py:97``)."""
    return np.asarray(points)[:int(number_of_points)]


def resample_batch(clouds: Sequence, number_of_points: int, mode: str = "random", first_stream: Optional[int] = None):
    """A batch of clouds -> one dense ``float32 [B, N, 3]`` device tensor (``DeviceArray``; exposes
    ``__cuda_array_interface__`` so ``torch.as_tensor(x, device="cuda")`` views it without a copy), the
    input layout of ``models/pointnet.py:65``.  ``mode``: ``"random"`` or ``"prefix"``."""
    arrs = [np.asarray(c.points if isinstance(c, PointCloud) else c, dtype=np.float32).reshape(-1, 3) for c in clouds]
    B, N = len(arrs), int(number_of_points)
    off = np.zeros(B + 1, np.int64)
    off[1:] = np.cumsum([a.shape[0] for a in arrs])
    if mode == "random" and any(a.shape[0] < N for a in arrs):
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    ctx = _cabi.default_context()
    flat = ctx.to_device(np.concatenate(arrs, axis=0) if B else np.zeros((0, 3), np.float32), np.float32)
    out = ctx.empty((B, N, 3), np.float32)
    counts = np.zeros(B, np.int64)
    if first_stream is None:
        first_stream = next(_state["stream"])
        for _ in range(max(B - 1, 0)):
            next(_state["stream"])
    m = _cabi.RESAMPLE_RANDOM if mode == "random" else _cabi.RESAMPLE_PREFIX
    ctx.check(ctx.lib.kp_resample_batch(ctx.handle, flat.ptr, off.ctypes.data_as(C.POINTER(C.c_int64)), B, N, m, _state["seed"],
                                        int(first_stream), out.ptr, counts.ctypes.data_as(C.POINTER(C.c_int64))))
    return out, counts


def statistical_outlier_removal(pcd: PointCloud, nb_neighbors: int = 200, std_ratio: float = 3.0) -> PointCloud:
    """``utils/processing.py:302-310``: 0.02 voxel, then ``remove_statistical_outlier``."""
    voxel_down_pcd = pcd.voxel_down_sample(voxel_size=0.02)
    filtered_cloud, _ = voxel_down_pcd.remove_statistical_outlier(nb_neighbors, std_ratio)
    return filtered_cloud
