"""On-disk formats either side of the hot path: ``.pcd`` clouds as written
by ``o3d.io.write_point_cloud`` (``preprocessing/data.py:69``, ``floor_removal.py:78``) and read by
``floor_removal.py:61`` / ``datasets/kinect_dataset.py:103``.  Host-side, numpy only."""
from __future__ import annotations

import numpy as np

from .geometry import PointCloud


def write_point_cloud(filename: str, pcd: PointCloud, write_ascii: bool = False) -> bool:
    pts = np.asarray(pcd.points, dtype=np.float32)
    n = pts.shape[0]
    has_c = pcd.has_colors()
    fields = "x y z" + (" rgb" if has_c else "")
    sizes = "4 4 4" + (" 4" if has_c else "")
    types = "F F F" + (" F" if has_c else "")
    counts = "1 1 1" + (" 1" if has_c else "")
    header = (f"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS {fields}\nSIZE {sizes}\nTYPE {types}\n"
              f"COUNT {counts}\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\n"
              f"DATA {'ascii' if write_ascii else 'binary'}\n")
    if has_c:
        c8 = np.clip(np.floor(np.asarray(pcd.colors) * 255.0), 0, 255).astype(np.uint32)
        packed = ((c8[:, 0] << 16) | (c8[:, 1] << 8) | c8[:, 2]).astype(np.uint32).view(np.float32)
        rec = np.empty(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<f4")])
        rec["rgb"] = packed
    else:
        rec = np.empty(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4")])
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        if write_ascii:
            for r in rec:
                f.write((" ".join(repr(float(v)) for v in r) + "\n").encode("ascii"))
        else:
            f.write(rec.tobytes())
    return True


def read_point_cloud(filename: str) -> PointCloud:
    with open(filename, "rb") as f:
        hdr = {}
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if not line or line.startswith("#"):
                if not line:
                    break
                continue
            k, _, v = line.partition(" ")
            hdr[k] = v
            if k == "DATA":
                break
        fields = hdr["FIELDS"].split()
        sizes = [int(s) for s in hdr["SIZE"].split()]
        types = hdr["TYPE"].split()
        n = int(hdr["POINTS"])
        np_t = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 4): "<u4", ("I", 4): "<i4", ("U", 1): "u1", ("U", 2): "<u2",
                ("I", 2): "<i2", ("I", 1): "i1"}
        dt = np.dtype([(fn, np_t[(t, s)]) for fn, t, s in zip(fields, types, sizes)])
        if hdr["DATA"] == "binary":
            rec = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
        elif hdr["DATA"] == "ascii":
            rows = np.loadtxt(f, ndmin=2) if n else np.zeros((0, len(fields)))
            rec = np.empty(n, dtype=dt)
            for j, fn in enumerate(fields):
                rec[fn] = rows[:, j]
        else:
            raise ValueError("binary_compressed .pcd is not supported")
    pcd = PointCloud(np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float64))
    if "rgb" in fields and n:
        raw = np.ascontiguousarray(rec["rgb"]).view(np.uint32)
        pcd.colors = np.stack([(raw >> 16) & 255, (raw >> 8) & 255, raw & 255], axis=1).astype(np.float64) / 255.0
    return pcd
