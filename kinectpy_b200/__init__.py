"""kinectpy_b200 -- B200-native per-frame point-cloud preprocessing path of KinectPy.

Call surface (same names, argument order and defaults as the reference):

* ``kinectpy_b200.preprocessing.filtering``      <- ``preprocessing/filtering.py``
* ``kinectpy_b200.preprocessing.registration``   <- ``preprocessing/registration.py``
* ``kinectpy_b200.floor_removal``                <- ``floor_removal.py``
* ``kinectpy_b200.utils.io``                     <- ``utils/io.py``
* ``kinectpy_b200.o3d``                          <- the slice of ``open3d`` those modules touch

Everything computes on the GPU through ``libkinectpy_b200.so`` (C ABI: ``include/kp_api.h``);
there is no CPU fallback.
"""
from . import _cabi
from ._cabi import KinectPyB200Error, device_available
from .geometry import PointCloud

__all__ = ["PointCloud", "KinectPyB200Error", "device_available"]
__version__ = "0.1.0"
