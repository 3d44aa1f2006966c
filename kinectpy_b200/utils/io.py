"""B200 counterpart of the reference's ``utils/io.py``: ``load_color``, ``load_depth``,
``rgbd_to_pointcloud`` with the same signatures.

``load_depth`` reads the external extractor's ``<ts>_depth.dat`` -- int16 XYZ triplets in
millimetres, shape ``(H*W, 3)`` (``utils/io.py:15-20``).  ``rgbd_to_pointcloud`` converts them to a
cloud, drops every pixel with ``x == 0 or y == 0 or z == 0`` (``utils/io.py:36``) and scales colours
to [0,1]; the conversion, the validity rule and the ordered compaction run on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _cabi
from .._cabi import KinectPyB200Error
from ..geometry import PointCloud

COLOR_SUFFIX = '_rgb.png'
DEPTH_SUFFIX = '_depth.dat'


def load_color(color_fp: str) -> np.ndarray:
    import cv2  # deferred: OpenCV is only needed to decode the PNG
    if not color_fp.endswith(COLOR_SUFFIX):
        color_fp += COLOR_SUFFIX
    return cv2.cvtColor(cv2.imread(color_fp), cv2.COLOR_BGR2RGB)


def load_depth(depth_fp: str) -> np.ndarray:
    if not depth_fp.endswith(DEPTH_SUFFIX):
        depth_fp += DEPTH_SUFFIX
    return np.fromfile(depth_fp, dtype=np.int16).reshape(-1, 3)


def save_depth(depth_fp: str, xyz16: np.ndarray) -> None:
    """Inverse of ``load_depth`` (same layout the k4a extractor writes)."""
    if not depth_fp.endswith(DEPTH_SUFFIX):
        depth_fp += DEPTH_SUFFIX
    np.ascontiguousarray(xyz16, dtype=np.int16).reshape(-1, 3).tofile(depth_fp)


def rgbd_to_pointcloud(color_img, depth_img, keep_mask=None, transform=None, _on_device=None) -> PointCloud:
    """int16 XYZ (+ RGB) -> cloud of the valid pixels, in pixel order (``utils/io.py:23-43``).

    ``keep_mask`` (uint8 per pixel) and ``transform`` (4x4) are extensions used by the crop / fusion
    callers so that the mask, the extrinsic and the compaction stay in one pass on the device.
    ``_on_device`` = ``(d_xyz16, d_keep)``: the crop caller's own device copies of ``depth_img`` and of the mask it just
    computed there (no trip of the mask to the host and back, no second upload of the frame).
    """
    raw = np.asarray(depth_img).reshape(-1, 3)
    if raw.dtype != np.int16:
        # the reference casts to float64 (utils/io.py:29); the device path is the int16 `_depth.dat` layout: anything that
        # is not exactly representable there is refused instead of being wrapped or truncated silently
        as16 = raw.astype(np.int16)
        if not np.array_equal(as16.astype(np.float64), raw.astype(np.float64)):
            raise KinectPyB200Error(_cabi.KP_E_ARG, "rgbd_to_pointcloud: depth_img holds values that are not int16 XYZ millimetres")
        raw = as16
    xyz16 = np.ascontiguousarray(raw, dtype=np.int16)
    n = xyz16.shape[0]
    if n == 0:
        return PointCloud()
    ctx = _cabi.default_context()
    if _on_device is not None:
        d16, d_keep = _on_device
    else:
        d16 = ctx.to_device(xyz16)
        d_keep = ctx.to_device(np.ascontiguousarray(keep_mask, dtype=np.uint8).reshape(-1)) if keep_mask is not None else None
    pts = ctx.empty((n, 3), np.float32)
    t16 = _cabi.T16(transform) if transform is not None else None
    ctx.check(ctx.lib.kp_points_from_xyz16(ctx.handle, d16.ptr, n, t16.ctypes.data if t16 is not None else None,
                                           _cabi.UNPROJECT_DROP_ANY_ZERO, 1.0, d_keep.ptr if d_keep is not None else None,
                                           pts.ptr, None))
    col = None
    if color_img is not None:
        c = np.asarray(color_img).astype(np.float64).reshape(-1, 3) / 255
        col = ctx.to_device(c, dtype=np.float32)
    o_p = ctx.empty((n, 3), np.float32)
    o_c = ctx.empty((n, 3), np.float32) if col is not None else None
    cnt = C.c_int64()
    ctx.check(ctx.lib.kp_compact(ctx.handle, n, None, 0, pts.ptr, o_p.ptr, col.ptr if col is not None else None,
                                 o_c.ptr if o_c is not None else None, None, None, None, C.byref(cnt)))
    return PointCloud._from_device(ctx, o_p, cnt.value, o_c, None)
