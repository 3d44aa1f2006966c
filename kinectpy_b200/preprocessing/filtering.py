"""B200 counterpart of the reference's ``preprocessing/filtering.py``.

``filter_outliers`` keeps the reference signature and defaults (``filtering.py:12-17``) and the same
composition -- voxel downsample, then statistical outlier removal, indices discarded
(``filtering.py:23-25``) -- but both steps run as CUDA kernels (radix-sort voxelisation, grid-hash
kNN).  ``Filtering`` (Mask R-CNN person segmentation, ``filtering.py:28-95``) and
``kalman_filter`` (``filtering.py:98-129``) are CPU-side helpers outside the point-cloud path
(SURVEY.md section 2, rows 3-4); they are provided so the module's surface stays complete.
"""
from __future__ import annotations

import copy

import numpy as np

from ..geometry import PointCloud


def filter_outliers(pcd: PointCloud, nb_neighbors: int = 200, std_ratio: float = 3.0,
                    voxel_size: float = 0.02) -> PointCloud:
    """Voxel-downsample a *copy* of ``pcd`` and drop statistical outliers; the input is untouched."""
    down = copy.deepcopy(pcd).voxel_down_sample(voxel_size)
    kept, _ = down.remove_statistical_outlier(nb_neighbors, std_ratio)
    return kept


class Filtering:
    """Person segmentation with an OpenCV-dnn Mask R-CNN (not a point-cloud op; CPU passthrough).

    Output contract (the seam into the crop of ``preprocessing/data.py:165-178``): an RGB image of the
    input's shape where every non-person pixel is black.
    """

    def __init__(self, frozen_graph_fp, pbtxt_fp):
        import cv2  # deferred: only this class needs OpenCV
        self._cv2 = cv2
        self.net = cv2.dnn.readNetFromTensorflow(frozen_graph_fp, pbtxt_fp)
        self.score_threshold = 0.5
        self.mask_threshold = 0.1

    def apply_segmentation(self, img: np.ndarray) -> np.ndarray:
        cv2 = self._cv2
        h, w = img.shape[:2]
        self.net.setInput(cv2.dnn.blobFromImage(img, swapRB=True))
        boxes, masks = self.net.forward(["detection_out_final", "detection_masks"])
        person = np.zeros((h, w), dtype=np.uint8)
        for det in boxes[0, 0]:
            cls, score = int(det[1]), float(det[2])
            if cls != 0 or score < self.score_threshold:
                continue
            x0, y0 = max(int(det[3] * w), 0), max(int(det[4] * h), 0)
            x1, y1 = min(int(det[5] * w), w - 1), min(int(det[6] * h), h - 1)
            if x1 <= x0 or y1 <= y0:
                continue
            m = cv2.resize(masks[int(np.where((boxes[0, 0] == det).all(axis=1))[0][0]), cls], (x1 - x0 + 1, y1 - y0 + 1))
            roi = (m > self.mask_threshold).astype(np.uint8)
            contours, _ = cv2.findContours(roi, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
            filled = np.zeros_like(roi)
            cv2.fillPoly(filled, contours, 1)
            person[y0:y1 + 1, x0:x1 + 1] |= filled
        out = img.copy()
        out[person == 0] = 0
        return out


def kalman_filter(joint_vals, ri=10, qi=10, fi=1 / 30, hi=1):
    """Constant-model Kalman smoother over an ``(N,3)`` joint trajectory (skeleton side, sequential)."""
    z = np.asarray(joint_vals, dtype=np.float64)
    n = z.shape[0]
    eye = np.eye(3)
    F, H, Q, R = fi * eye, hi * eye, qi * eye, ri * eye
    x = np.zeros((n, 3))
    P = eye.copy()
    if n:
        x[0] = z[0]
    for k in range(1, n):
        x_pred = F @ x[k - 1]
        P_pred = F @ P @ F.T + Q
        K = P_pred @ H.T @ np.linalg.inv(H @ P_pred @ H.T + R)
        x[k] = x_pred + K @ (z[k] - H @ x_pred)
        P = (eye - K @ H) @ P_pred
    return x
