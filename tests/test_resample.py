"""K6 fixed-N resampling (BASELINE config C5): select_points_randomly (utils/processing.py:259-275) and the
points[:N] prefix (datasets/kinect_dataset_npz.py:97).  CPU tests pin the oracle; GPU tests compare the kernels
with it bit for bit (indices and coordinates)."""
import numpy as np
import pytest


def _cloud(n, seed):
    r = np.random.default_rng(seed)
    return (r.random((n, 3)) * 4 - 2).astype(np.float32)


# ------------------------------------------------------------------ oracle (CPU)
def test_oracle_resample_keys_match_scalar_rng(oracle):
    k = oracle.rng_array(1234, 3, np.arange(64))
    assert [int(v) for v in k] == [oracle.rng(1234, 3, i) for i in range(64)]


def test_oracle_resample_is_a_distinct_uniform_subset(oracle):
    pts = _cloud(5000, 1)
    pts[::7] = np.nan
    sel, idx = oracle.resample_fixed_n(pts, 1024, seed=9, stream=4)
    assert idx.shape == (1024,) and len(set(idx.tolist())) == 1024           # replace=False
    assert not np.isnan(sel).any() and np.array_equal(sel, pts[idx])
    # a different stream / seed gives a different subset; the same one is reproducible
    assert not np.array_equal(idx, oracle.resample_fixed_n(pts, 1024, seed=9, stream=5)[1])
    assert np.array_equal(idx, oracle.resample_fixed_n(pts, 1024, seed=9, stream=4)[1])
    # uniformity: over many streams every valid point is picked about N/n_valid of the time
    n = 400
    small = _cloud(n, 2)
    hits = np.zeros(n)
    for s in range(300):
        hits[oracle.resample_fixed_n(small, 100, seed=1, stream=s)[1]] += 1
    assert abs(hits.mean() - 75.0) < 1e-9 and hits.std() < 12.0               # binomial(300, .25): sigma = 7.5
    # order is random too: first element is not biased towards low indices
    firsts = [oracle.resample_fixed_n(small, 10, seed=2, stream=s)[1][0] for s in range(200)]
    assert 150 < np.mean(firsts) < 250
    with pytest.raises(ValueError):
        oracle.resample_fixed_n(pts, 5000)                                    # NaN rows are not selectable
    p, i = oracle.resample_fixed_n(small, 1000, mode="prefix")
    assert p.shape == (400, 3) and np.array_equal(p, small)
    p, i = oracle.resample_fixed_n(small, 100, mode="prefix")
    assert np.array_equal(p, small[:100])


def test_processing_mirror_signatures():
    import inspect
    from kinectpy_b200.utils import processing as P
    sig = inspect.signature(P.select_points_randomly)
    assert list(sig.parameters)[:2] == ["pointcloud", "number_of_points"]     # utils/processing.py:259-262
    sig = inspect.signature(P.statistical_outlier_removal)
    assert [p.default for p in sig.parameters.values()][1:] == [200, 3.0]     # utils/processing.py:302


# ------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("n,N", [(100000, 4096), (4096, 4096), (5000, 1), (640000, 4096)])
def test_gpu_resample_bit_exact(ctx, oracle, n, N):
    import gpu_helpers as G
    pts = _cloud(n, n + N)
    ref_p, ref_i = oracle.resample_fixed_n(pts, N, seed=77, stream=3)
    got_p, got_i = G.resample(ctx, pts, N, 0, 77, 3)
    assert np.array_equal(got_i, ref_i) and np.array_equal(got_p, ref_p)


@pytest.mark.gpu
def test_gpu_resample_edges(ctx, oracle):
    import gpu_helpers as G
    from kinectpy_b200 import KinectPyB200Error
    pts = _cloud(3000, 5)
    pts[::3] = np.nan
    ref_p, ref_i = oracle.resample_fixed_n(pts, 2000, seed=1, stream=0)
    got_p, got_i = G.resample(ctx, pts, 2000, 0, 1, 0)
    assert np.array_equal(got_i, ref_i) and np.array_equal(got_p, ref_p)
    with pytest.raises(KinectPyB200Error):
        G.resample(ctx, pts, 2001, 0, 1, 0)          # only 2000 valid points
    with pytest.raises(KinectPyB200Error):
        G.resample(ctx, pts, 3001, 0, 1, 0)          # larger than the population
    p, i = G.resample(ctx, pts, 0, 0, 1, 0)
    assert p.shape[0] == 0
    p, _ = G.resample(ctx, pts, 5000, 1)             # prefix on a short cloud
    assert p.shape[0] == 3000 and np.array_equal(p, pts, equal_nan=True)
    p, _ = G.resample(ctx, pts, 10, 1)
    assert np.array_equal(p, pts[:10], equal_nan=True)


@pytest.mark.gpu
def test_gpu_resample_batch_and_surface(ctx, oracle):
    import gpu_helpers as G
    clouds = [_cloud(n, 40 + n) for n in (6000, 4500, 9000)]
    out, counts = G.resample_batch(ctx, clouds, 4096, 0, 5, 10)
    assert out.shape == (3, 4096, 3) and counts.tolist() == [4096] * 3
    for b, c in enumerate(clouds):
        assert np.array_equal(out[b], oracle.resample_fixed_n(c, 4096, seed=5, stream=10 + b)[0])
    out, counts = G.resample_batch(ctx, clouds, 5000, 1)
    assert counts.tolist() == [5000, 4500, 5000]
    assert np.array_equal(out[1, :4500], clouds[1]) and not out[1, 4500:].any()       # zero padded
    # the reference-facing call surface
    from kinectpy_b200.geometry import PointCloud
    from kinectpy_b200.utils import processing as P
    P.seed(5)
    pcd = PointCloud(clouds[0].astype(np.float64))
    sel = P.select_points_randomly(pcd, 4096, stream=10)
    assert sel.dtype == np.float64 and np.array_equal(sel.astype(np.float32), out[0] if False else oracle.resample_fixed_n(clouds[0], 4096, seed=5, stream=10)[0])
    with pytest.raises(ValueError):
        P.select_points_randomly(pcd, 6001)
    dev, cnt = P.resample_batch([pcd, PointCloud(clouds[2].astype(np.float64))], 4096, first_stream=0)
    assert dev.shape == (2, 4096, 3) and dev.to_host().shape == (2, 4096, 3)
