import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        from kinectpy_b200 import _cabi
        return _cabi.device_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def ctx():
    from kinectpy_b200 import _cabi
    return _cabi.default_context()


def make_surface_cloud(n, seed=0, scale=1.0, outliers=0.01, quantum=None):
    """Points on a few smooth sheets + isolated outliers: the shape of a fused Kinect cloud."""
    r = np.random.default_rng(seed)
    u = r.uniform(-1, 1, (n, 2))
    sheet = r.integers(0, 3, n)
    z = np.where(sheet == 0, 0.2 * np.sin(3 * u[:, 0]) + 0.1 * u[:, 1],
                 np.where(sheet == 1, 1.0 + 0.05 * u[:, 0] ** 2, -0.6 + 0.3 * u[:, 0] * u[:, 1]))
    pts = np.stack([u[:, 0], u[:, 1], z], axis=1) + r.normal(0, 0.002, (n, 3))
    k = int(n * outliers)
    if k:
        pts[r.choice(n, k, replace=False)] = r.uniform(-1.5, 1.5, (k, 3))
    pts *= scale
    if quantum:
        pts = np.round(pts / quantum) * quantum
    return pts.astype(np.float32)
