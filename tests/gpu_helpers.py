"""Thin numpy-in / numpy-out wrappers over the C ABI for the GPU parity tests: host arrays are
copied to the device, the entry point of ``include/kp_api.h`` is called through ctypes, results come
back as numpy.  (The product's Python surface is tested separately in test_gpu_surface.py.)"""
import ctypes as C

import numpy as np

from kinectpy_b200 import _cabi


def _ptr(d):
    return None if d is None else d.ptr


def unproject(ctx, depth, tab, T, flags, scale, want=("valid", "xyz16", "bounds")):
    B, S, P = depth.shape
    d_depth = ctx.to_device(depth, np.uint16)
    d_tab = ctx.to_device(tab, np.float32)
    Tc = None if T is None else np.ascontiguousarray(T, dtype=np.float64).reshape(S * 16)
    xyz = ctx.empty((B, S * P, 3), np.float32)
    valid = ctx.empty((B, S * P), np.uint8) if "valid" in want else None
    xyz16 = ctx.empty((B, S * P, 3), np.int16) if "xyz16" in want and (flags & 1) else None
    bounds = ctx.empty((B, 6), np.float32) if "bounds" in want else None
    nvalid = ctx.empty((B,), np.int32) if "bounds" in want else None
    ctx.check(ctx.lib.kp_unproject_transform(ctx.handle, d_depth.ptr, d_tab.ptr, None if Tc is None else Tc.ctypes.data,
                                             B, S, P, flags, float(scale), xyz.ptr, _ptr(valid), _ptr(xyz16),
                                             _ptr(bounds), _ptr(nvalid)))
    out = {"xyz": xyz.to_host()}
    for k, v in (("valid", valid), ("xyz16", xyz16), ("bounds", bounds), ("nvalid", nvalid)):
        out[k] = None if v is None else v.to_host()
    return out


def points_from_xyz16(ctx, xyz16, T=None, flags=2, scale=1.0, keep=None):
    a = np.ascontiguousarray(xyz16, np.int16).reshape(-1, 3)
    n = a.shape[0]
    d = ctx.to_device(a)
    dk = None if keep is None else ctx.to_device(np.ascontiguousarray(keep, np.uint8))
    xyz = ctx.empty((n, 3), np.float32)
    valid = ctx.empty((n,), np.uint8)
    t = None if T is None else _cabi.T16(T)
    ctx.check(ctx.lib.kp_points_from_xyz16(ctx.handle, d.ptr, n, None if t is None else t.ctypes.data, flags, float(scale),
                                           _ptr(dk), xyz.ptr, valid.ptr))
    return xyz.to_host(), valid.to_host()


def crop_mask(ctx, rgb, xyz16, gate):
    rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
    a = np.ascontiguousarray(xyz16, np.int16).reshape(-1, 3)
    n = a.shape[0]
    keep = ctx.empty((n,), np.uint8)
    med = C.c_double()
    d_rgb, d_a = ctx.to_device(rgb), ctx.to_device(a)   # keep the buffers alive across the call
    ctx.check(ctx.lib.kp_crop_mask(ctx.handle, d_rgb.ptr, d_a.ptr, n, float(gate), keep.ptr, C.byref(med)))
    return keep.to_host(), med.value


def transform(ctx, xyz, T, rotate_only=False):
    d = ctx.to_device(xyz, np.float32)
    t = _cabi.T16(T)
    ctx.check(ctx.lib.kp_transform_points(ctx.handle, d.ptr, d.shape[0], t.ctypes.data, 1 if rotate_only else 0))
    return d.to_host()


def bounds(ctx, xyz):
    d = ctx.to_device(xyz, np.float32)
    b = (C.c_float * 6)()
    nv = C.c_int64()
    ctx.check(ctx.lib.kp_bounds(ctx.handle, d.ptr, d.shape[0], b, C.byref(nv)))
    return np.array(list(b), np.float32), nv.value


def compact(ctx, xyz, mask=None, invert=False, colors=None):
    d = ctx.to_device(xyz, np.float32)
    n = d.shape[0]
    dm = None if mask is None else ctx.to_device(mask, np.uint8)
    dc = None if colors is None else ctx.to_device(colors, np.float32)
    o = ctx.empty((n, 3), np.float32)
    oc = None if colors is None else ctx.empty((n, 3), np.float32)
    idx = ctx.empty((n,), np.int32)
    cnt = C.c_int64()
    ctx.check(ctx.lib.kp_compact(ctx.handle, n, _ptr(dm), 1 if invert else 0, d.ptr, o.ptr, _ptr(dc), _ptr(oc), None, None,
                                 idx.ptr, C.byref(cnt)))
    m = cnt.value
    return o.to_host(m), idx.to_host(m), (None if oc is None else oc.to_host(m))


def voxel(ctx, xyz, voxel_size, colors=None, normals=None):
    d = ctx.to_device(xyz, np.float32)
    n = d.shape[0]
    dc = None if colors is None else ctx.to_device(colors, np.float32)
    dn = None if normals is None else ctx.to_device(normals, np.float32)
    o = ctx.empty((max(n, 1), 3), np.float32)
    oc = None if colors is None else ctx.empty((max(n, 1), 3), np.float32)
    on = None if normals is None else ctx.empty((max(n, 1), 3), np.float32)
    ijk = ctx.empty((max(n, 1), 3), np.int32)
    pv = ctx.empty((max(n, 1),), np.int32)
    minb = (C.c_double * 3)()
    m = C.c_int64()
    ctx.check(ctx.lib.kp_voxel_downsample(ctx.handle, d.ptr, _ptr(dc), _ptr(dn), n, float(voxel_size), o.ptr, _ptr(oc),
                                          _ptr(on), ijk.ptr, pv.ptr, minb, C.byref(m)))
    m = m.value
    return {"points": o.to_host(m), "colors": None if oc is None else oc.to_host(m),
            "normals": None if on is None else on.to_host(m), "ijk": ijk.to_host(m), "point_voxel": pv.to_host(n),
            "min_bound": np.array(list(minb)), "m": m}


def knn(ctx, xyz, k, queries=None, radius=0.0, cell_hint=0.0):
    d = ctx.to_device(xyz, np.float32)
    n = d.shape[0]
    dq = None if queries is None else ctx.to_device(queries, np.float32)
    nq = n if dq is None else dq.shape[0]
    idx = ctx.empty((nq, k), np.int32)
    d2 = ctx.empty((nq, k), np.float64)
    cnt = ctx.empty((nq,), np.int32)
    ctx.check(ctx.lib.kp_knn(ctx.handle, d.ptr, n, _ptr(dq), nq, k, float(radius), float(cell_hint), idx.ptr, d2.ptr, cnt.ptr))
    return idx.to_host(), d2.to_host(), cnt.to_host()


def sor(ctx, xyz, k, ratio, cell_hint=0.0):
    d = ctx.to_device(xyz, np.float32)
    n = d.shape[0]
    keep = ctx.empty((n,), np.uint8)
    mean = ctx.empty((n,), np.float64)
    stats = (C.c_double * 3)()
    kept = C.c_int64()
    ctx.check(ctx.lib.kp_sor_mask(ctx.handle, d.ptr, n, k, float(ratio), float(cell_hint), keep.ptr, mean.ptr, stats,
                                  C.byref(kept)))
    return keep.to_host(), mean.to_host(), np.array(list(stats)), kept.value


def radius_outlier(ctx, xyz, nb, radius):
    d = ctx.to_device(xyz, np.float32)
    n = d.shape[0]
    keep = ctx.empty((n,), np.uint8)
    cnt = ctx.empty((n,), np.int32)
    kept = C.c_int64()
    ctx.check(ctx.lib.kp_radius_mask(ctx.handle, d.ptr, n, nb, float(radius), keep.ptr, cnt.ptr, C.byref(kept)))
    return keep.to_host(), cnt.to_host(), kept.value


def normals(ctx, xyz, radius, max_nn):
    d = ctx.to_device(xyz, np.float32)
    out = ctx.empty(d.shape, np.float32)
    ctx.check(ctx.lib.kp_estimate_normals(ctx.handle, d.ptr, d.shape[0], float(radius), max_nn, out.ptr))
    return out.to_host()


def ransac(ctx, xyz, thr, ransac_n, iters, probability=0.99999999, seed=1234):
    d = ctx.to_device(xyz, np.float32)
    n = d.shape[0]
    mask = ctx.empty((n,), np.uint8)
    counts = ctx.empty((iters,), np.int64)
    plane = (C.c_double * 4)()
    ninl, best = C.c_int64(), C.c_int32()
    ctx.check(ctx.lib.kp_ransac_plane(ctx.handle, d.ptr, n, float(thr), ransac_n, iters, float(probability), seed, plane,
                                      mask.ptr, C.byref(ninl), C.byref(best), counts.ptr))
    return np.array(list(plane)), mask.to_host(), best.value, counts.to_host(), ninl.value


def plane_side(ctx, xyz, a, b, c, d_):
    d = ctx.to_device(xyz, np.float32)
    mask = ctx.empty((d.shape[0],), np.uint8)
    kept = C.c_int64()
    ctx.check(ctx.lib.kp_plane_side_mask(ctx.handle, d.ptr, d.shape[0], a, b, c, d_, mask.ptr, C.byref(kept)))
    return mask.to_host(), kept.value


def band(ctx, xyz, band_, axis=1):
    d = ctx.to_device(xyz, np.float32)
    mask = ctx.empty((d.shape[0],), np.uint8)
    amax, nlow = C.c_double(), C.c_int64()
    ctx.check(ctx.lib.kp_band_mask(ctx.handle, d.ptr, d.shape[0], axis, float(band_), mask.ptr, C.byref(amax), C.byref(nlow)))
    return mask.to_host(), amax.value, nlow.value


def icp(ctx, src, tgt, tgt_n, max_corr, init=None, max_iter=30, rel_fit=1e-6, rel_rmse=1e-6):
    ds, dt, dn = ctx.to_device(src, np.float32), ctx.to_device(tgt, np.float32), ctx.to_device(tgt_n, np.float32)
    T0 = _cabi.T16(np.eye(4) if init is None else init)
    T = np.zeros(16)
    fit, rmse, iters, nc = C.c_double(), C.c_double(), C.c_int(), C.c_int64()
    ctx.check(ctx.lib.kp_icp_point_to_plane(ctx.handle, ds.ptr, ds.shape[0], dt.ptr, dn.ptr, dt.shape[0], float(max_corr),
                                            T0.ctypes.data, max_iter, rel_fit, rel_rmse, T.ctypes.data, C.byref(fit),
                                            C.byref(rmse), C.byref(iters), C.byref(nc)))
    return {"T": T.reshape(4, 4), "fitness": fit.value, "rmse": rmse.value, "iters": iters.value, "ncorr": nc.value}


def resample(ctx, xyz, N, mode=0, seed=1234, stream=0):
    d = ctx.to_device(xyz, np.float32)
    n = d.shape[0]
    out = ctx.empty((max(N, 1), 3), np.float32)
    idx = ctx.empty((max(N, 1),), np.int32)
    cnt = C.c_int64()
    ctx.check(ctx.lib.kp_resample_fixed_n(ctx.handle, d.ptr, n, N, mode, seed, stream, out.ptr, idx.ptr, C.byref(cnt)))
    m = cnt.value
    return out.to_host(m), idx.to_host(m)


def resample_batch(ctx, clouds, N, mode=0, seed=1234, first_stream=0):
    off = np.zeros(len(clouds) + 1, np.int64)
    off[1:] = np.cumsum([c.shape[0] for c in clouds])
    flat = ctx.to_device(np.concatenate(clouds, axis=0), np.float32)
    out = ctx.empty((len(clouds), N, 3), np.float32)
    counts = np.zeros(len(clouds), np.int64)
    ctx.check(ctx.lib.kp_resample_batch(ctx.handle, flat.ptr, off.ctypes.data_as(C.POINTER(C.c_int64)), len(clouds), N, mode,
                                        seed, first_stream, out.ptr, counts.ctypes.data_as(C.POINTER(C.c_int64))))
    return out.to_host(), counts


def icp_p2p(ctx, src, tgt, max_corr, init=None, max_iter=30, rel_fit=1e-6, rel_rmse=1e-6):
    ds, dt = ctx.to_device(src, np.float32), ctx.to_device(tgt, np.float32)
    T0 = _cabi.T16(np.eye(4) if init is None else init)
    T = np.zeros(16)
    fit, rmse, iters, nc = C.c_double(), C.c_double(), C.c_int(), C.c_int64()
    ctx.check(ctx.lib.kp_icp_point_to_point(ctx.handle, ds.ptr, ds.shape[0], dt.ptr, dt.shape[0], float(max_corr),
                                            T0.ctypes.data, max_iter, rel_fit, rel_rmse, T.ctypes.data, C.byref(fit),
                                            C.byref(rmse), C.byref(iters), C.byref(nc)))
    return {"T": T.reshape(4, 4), "fitness": fit.value, "rmse": rmse.value, "iters": iters.value, "ncorr": nc.value}


def icp_colored(ctx, src, src_col, tgt, tgt_col, tgt_n, max_corr, init=None, max_iter=30, rel_fit=1e-6, rel_rmse=1e-6, lam=0.968):
    ds, dsc = ctx.to_device(src, np.float32), ctx.to_device(src_col, np.float32)
    dt, dtc, dn = ctx.to_device(tgt, np.float32), ctx.to_device(tgt_col, np.float32), ctx.to_device(tgt_n, np.float32)
    T0 = _cabi.T16(np.eye(4) if init is None else init)
    T = np.zeros(16)
    fit, rmse, iters, nc = C.c_double(), C.c_double(), C.c_int(), C.c_int64()
    ctx.check(ctx.lib.kp_icp_colored(ctx.handle, ds.ptr, dsc.ptr, ds.shape[0], dt.ptr, dtc.ptr, dn.ptr, dt.shape[0],
                                     float(max_corr), float(lam), T0.ctypes.data, max_iter, rel_fit, rel_rmse, T.ctypes.data,
                                     C.byref(fit), C.byref(rmse), C.byref(iters), C.byref(nc)))
    return {"T": T.reshape(4, 4), "fitness": fit.value, "rmse": rmse.value, "iters": iters.value, "ncorr": nc.value}


def color_gradient(ctx, xyz, colors, normals, radius, max_nn=30):
    d, dc, dn = ctx.to_device(xyz, np.float32), ctx.to_device(colors, np.float32), ctx.to_device(normals, np.float32)
    out = ctx.empty(d.shape, np.float32)
    ctx.check(ctx.lib.kp_color_gradient(ctx.handle, d.ptr, dc.ptr, dn.ptr, d.shape[0], float(radius), max_nn, out.ptr))
    return out.to_host()


def fpfh(ctx, xyz, normals, radius, max_nn):
    d, dn = ctx.to_device(xyz, np.float32), ctx.to_device(normals, np.float32)
    out = ctx.empty((d.shape[0], 33), np.float64)
    ctx.check(ctx.lib.kp_fpfh(ctx.handle, d.ptr, dn.ptr, d.shape[0], float(radius), max_nn, out.ptr))
    return out.to_host()


def feature_match(ctx, fa, fb):
    da, db = ctx.to_device(fa, np.float64), ctx.to_device(fb, np.float64)
    nn = ctx.empty((da.shape[0],), np.int32)
    d2 = ctx.empty((da.shape[0],), np.float64)
    ctx.check(ctx.lib.kp_feature_match(ctx.handle, da.ptr, da.shape[0], db.ptr, db.shape[0], 33, nn.ptr, d2.ptr))
    return nn.to_host(), d2.to_host()


def ransac_correspondence(ctx, src, tgt, corres, max_corr, ransac_n=3, edge_sim=0.95, dist_thr=None, max_iter=250000,
                          confidence=0.999, seed=1234):
    ds, dt = ctx.to_device(src, np.float32), ctx.to_device(tgt, np.float32)
    dc = ctx.to_device(np.ascontiguousarray(corres, np.int32).reshape(-1, 2), np.int32)
    T = np.zeros(16)
    fit, rmse, best, val = C.c_double(), C.c_double(), C.c_int32(), C.c_int64()
    ctx.check(ctx.lib.kp_ransac_correspondence(ctx.handle, ds.ptr, ds.shape[0], dt.ptr, dt.shape[0], dc.ptr, dc.shape[0],
                                               float(max_corr), ransac_n, float(edge_sim), float(max_corr if dist_thr is None else dist_thr),
                                               max_iter, float(confidence), seed, T.ctypes.data_as(C.POINTER(C.c_double)),
                                               C.byref(fit), C.byref(rmse), C.byref(best), C.byref(val)))
    return {"T": T.reshape(4, 4), "fitness": fit.value, "rmse": rmse.value, "best_iter": best.value, "validated": val.value}
