"""B200 counterpart of the reference's ``preprocessing/registration.py``.

Same function names, argument order and defaults.  The voxel downsample, normal estimation and the
whole ICP loop (point-to-plane, point-to-point, coloured) run on the GPU.

Roles, as in the reference: ``prepare_dataset`` makes the *sub* cloud the ICP source and the
*master* cloud the target (``registration.py:25-26``), so the returned 4x4 maps sub coordinates
into the master frame; ``execute_point_to_plane_registration`` swaps its local names before
calling it (``registration.py:73-77``), which this module reproduces.
"""
from __future__ import annotations

import copy

import numpy as np

from .. import geometry as _g
from ..geometry import PointCloud


def preprocess_point_cloud(pcd: PointCloud, voxel_size, normals_nn=30, fpfh_nn=100):
    """voxel -> normals (hybrid search: radius 2*voxel, at most ``normals_nn``) (``registration.py:7-21``).

    then FPFH over the hybrid neighbourhood of radius 5*voxel, at most ``fpfh_nn``.  ``fpfh_nn=None`` skips
    the feature (the ICP wrapper throws it away, ``registration.py:77``)."""
    down = pcd.voxel_down_sample(voxel_size)
    down.estimate_normals(_g.KDTreeSearchParamHybrid(radius=voxel_size * 2, max_nn=normals_nn))
    if fpfh_nn is None:
        return down, None
    fpfh = _g.compute_fpfh_feature(down, _g.KDTreeSearchParamHybrid(radius=voxel_size * 5, max_nn=fpfh_nn))
    return down, fpfh


def prepare_dataset(pcd_master, pcd_sub, voxel_size, normals_nn=40, fpfh_nn=40):
    source, target = copy.deepcopy(pcd_sub), copy.deepcopy(pcd_master)
    source_down, source_fpfh = preprocess_point_cloud(source, voxel_size, normals_nn, fpfh_nn)
    target_down, target_fpfh = preprocess_point_cloud(target, voxel_size, normals_nn, fpfh_nn)
    return source, target, source_down, target_down, source_fpfh, target_fpfh


def execute_global_registration(pcd_master, pcd_sub, voxel_size: int = 35, ransac_n_trials: int = 15) -> np.ndarray:
    """FPFH + feature-matching RANSAC (``registration.py:32-62``): ``ransac_n_trials`` independent RANSAC runs
    (250 000 hypotheses, confidence 0.999, edge-length 0.95 and distance checkers, mutual filter), the
    transformation of the fittest one is returned; it maps ``pcd_sub`` into the frame of ``pcd_master``.
    ``prepare_dataset`` is deterministic, so it is evaluated once instead of once per trial."""
    best_fitness = 0
    ransac_transformation = None
    (source, target, source_down, target_down, source_fpfh, target_fpfh) = prepare_dataset(pcd_master, pcd_sub, voxel_size)
    distance_threshold = voxel_size * 1.5
    for _ in range(ransac_n_trials):
        result_ransac = _g.registration_ransac_based_on_feature_matching(
            source_down, target_down, source_fpfh, target_fpfh, True, distance_threshold,
            _g.TransformationEstimationPointToPoint(False), 3,
            [_g.CorrespondenceCheckerBasedOnEdgeLength(0.95), _g.CorrespondenceCheckerBasedOnDistance(distance_threshold)],
            _g.RANSACConvergenceCriteria(250000, 0.999))
        if best_fitness < result_ransac.fitness:
            best_fitness = result_ransac.fitness
            ransac_transformation = result_ransac.transformation
    return ransac_transformation


def execute_point_to_plane_registration(pcd_master, pcd_sub, initial_transformation, voxel_size: int = 35,
                                        threshold: float = 100, return_result: bool = False):
    """Point-to-plane ICP refinement of ``initial_transformation`` (``registration.py:65-86``).

    ``threshold`` is the literal 100 of ``registration.py:75`` exposed as a keyword (the reference
    works in millimetres; BASELINE configs use metres).  Open3D's default criteria apply:
    30 iterations, 1e-6 / 1e-6.
    """
    first, second = copy.deepcopy(pcd_master), copy.deepcopy(pcd_sub)
    # the reference passes (master, sub) into prepare_dataset(pcd_master, pcd_sub): source <- sub, target <- master.
    # Its FPFH features are computed and dropped there (registration.py:77); fpfh_nn=None skips them.
    _, _, source_down, target_down, _, _ = prepare_dataset(first, second, voxel_size, fpfh_nn=None)
    result = _g.registration_icp(source_down, target_down, threshold, initial_transformation,
                                 _g.TransformationEstimationPointToPlane())
    return result if return_result else result.transformation


def execute_colored_ICP_registration(pcd_master, pcd_sub, initial_transformation,
                                     voxel_radius=(80, 40, 20), max_iter=(50, 30, 14)):
    """Multi-scale coloured ICP (``registration.py:89-114``).  Restated as written: source <- master,
    target <- sub (``:92-93``), every scale starts again from ``initial_transformation`` (the reference
    never feeds a scale's result into the next), and the last scale's transform is
    returned.  ``voxel_radius`` / ``max_iter`` are the literals of ``:95-96`` exposed as keywords."""
    source, target = copy.deepcopy(pcd_master), copy.deepcopy(pcd_sub)
    result_icp = None
    for radius, iters in zip(voxel_radius, max_iter):
        source_down = source.voxel_down_sample(radius)
        target_down = target.voxel_down_sample(radius)
        source_down.estimate_normals(_g.KDTreeSearchParamHybrid(radius=radius * 2, max_nn=30))
        target_down.estimate_normals(_g.KDTreeSearchParamHybrid(radius=radius * 2, max_nn=30))
        result_icp = _g.registration_colored_icp(
            source_down, target_down, radius, initial_transformation, _g.TransformationEstimationForColoredICP(),
            _g.ICPConvergenceCriteria(relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=iters))
    return result_icp.transformation
