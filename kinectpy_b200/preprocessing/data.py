"""B200 counterpart of the per-frame helpers of the reference's ``preprocessing/data.py``.

The reference's ``DataProcessor`` is file / pandas scaffolding around three point-cloud steps: the human
crop (``data.py:165-178``), the per-sub ``transform`` (``data.py:46-48``) and the ``np.vstack`` fusion
followed by ``filter_outliers`` (``data.py:51-61``).  Those three steps are provided here as functions over
device-backed clouds; the filename table and the Mask R-CNN call stay with the caller.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from .. import _cabi
from ..geometry import PointCloud
from ..utils.io import rgbd_to_pointcloud
from .filtering import filter_outliers


def transform_filtered_image_to_pointcloud(filtered_img, depth_img, gate: float = 750, transform=None) -> PointCloud:
    """``DataProcessor._transform_filtered_image_to_pointcloud`` (``data.py:165-178``): keep pixels that are
    non-black in all three channels and whose ``z <= median(z) + gate`` (median over ALL pixels, zeros
    included; the reference's second OR-term is subsumed), then ``rgbd_to_pointcloud``.  Median (65 536-bin
    histogram), mask, validity rule, optional extrinsic and compaction run on the GPU."""
    rgb = np.ascontiguousarray(np.asarray(filtered_img).reshape(-1, 3), dtype=np.uint8)
    xyz16 = np.ascontiguousarray(np.asarray(depth_img).reshape(-1, 3), dtype=np.int16)
    n = xyz16.shape[0]
    if n == 0:
        return PointCloud()
    ctx = _cabi.default_context()
    d_rgb, d_xyz = ctx.to_device(rgb), ctx.to_device(xyz16)
    keep = ctx.empty((n,), np.uint8)
    ctx.check(ctx.lib.kp_crop_mask(ctx.handle, d_rgb.ptr, d_xyz.ptr, n, float(gate), keep.ptr, None))
    return rgbd_to_pointcloud(rgb, xyz16, transform=transform, _on_device=(d_xyz, keep))


def fuse_registered(clouds: Sequence[PointCloud], transforms: Sequence[np.ndarray]) -> PointCloud:
    """``data.py:44-58``: cloud 0 (master) as is, cloud i > 0 transformed in place by ``transforms[i-1]``,
    then concatenated in device order (master, sub_1, sub_2, ...)."""
    fused = None
    for i, pcd in enumerate(clouds):
        if i > 0:
            pcd.transform(transforms[i - 1])
        fused = pcd if fused is None else fused + pcd
    return fused if fused is not None else PointCloud()


def fuse_and_filter(clouds: Sequence[PointCloud], transforms: Sequence[np.ndarray], **filter_kwargs) -> PointCloud:
    """``data.py:44-61``: fuse, then ``filter_outliers`` with the reference defaults unless overridden."""
    return filter_outliers(fuse_registered(clouds, transforms), **filter_kwargs)
