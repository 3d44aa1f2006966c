"""The slice of the ``open3d`` namespace the reference's hot-path modules use, mapped onto the
B200 implementation.  ``import kinectpy_b200.o3d as o3d`` is the one-line switch for
``preprocessing/filtering.py``, ``preprocessing/registration.py``, ``floor_removal.py`` and
``utils/io.py``."""
from types import SimpleNamespace

from . import geometry as _g
from . import io_formats as _io


def _not_on_path(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name} is outside the B200 hot path (SURVEY.md section 8f)")
    return f


geometry = SimpleNamespace(
    PointCloud=_g.PointCloud,
    KDTreeSearchParamHybrid=_g.KDTreeSearchParamHybrid,
    KDTreeSearchParamKNN=_g.KDTreeSearchParamKNN,
)
utility = SimpleNamespace(Vector3dVector=_g.Vector3dVector, Vector2iVector=_g.Vector2iVector, random=SimpleNamespace(seed=_g.seed))
pipelines = SimpleNamespace(registration=SimpleNamespace(
    registration_icp=_g.registration_icp,
    TransformationEstimationPointToPlane=_g.TransformationEstimationPointToPlane,
    TransformationEstimationPointToPoint=_g.TransformationEstimationPointToPoint,
    ICPConvergenceCriteria=_g.ICPConvergenceCriteria,
    RegistrationResult=_g.RegistrationResult,
    compute_fpfh_feature=_g.compute_fpfh_feature,
    registration_ransac_based_on_feature_matching=_g.registration_ransac_based_on_feature_matching,
    Feature=_g.Feature,
    CorrespondenceCheckerBasedOnEdgeLength=_g.CorrespondenceCheckerBasedOnEdgeLength,
    CorrespondenceCheckerBasedOnDistance=_g.CorrespondenceCheckerBasedOnDistance,
    RANSACConvergenceCriteria=_g.RANSACConvergenceCriteria,
    registration_colored_icp=_g.registration_colored_icp,
    TransformationEstimationForColoredICP=_g.TransformationEstimationForColoredICP,
))
io = SimpleNamespace(read_point_cloud=_io.read_point_cloud, write_point_cloud=_io.write_point_cloud)
