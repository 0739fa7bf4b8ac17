"""PCD (Point Cloud Data, v0.7) reader / writer for pcl::PointXYZRGBA clouds.

The data format either side of the path: the model builder writes one `models/<time>/<k>.pcd` per cluster with
pcl::PCDWriter::write(..., binary=false) (ref: src/create_model.cpp:209-230) and offline runs of the tracker replay
frames stored as PCD files (BASELINE.json configs[0]).  Host-side I/O only: the points come back as the packed
{x, y, z, rgba} records that PointCloud(...) uploads.

Supported: DATA ascii and DATA binary; fields x, y, z (float32) plus an optional colour field -- `rgba` (uint32, what
PCL >= 1.7 writes for PointXYZRGBA) or `rgb` (the same 4 bytes stored as a float32); any other field is skipped.
`binary_compressed` is rejected.  Organised clouds (HEIGHT > 1) keep their row-major order; NaNs are kept.
"""
import numpy as np

from ._capi import POINT

_NP = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "<u1", ("U", 2): "<u2", ("U", 4): "<u4", ("U", 8): "<u8",
       ("I", 1): "<i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8"}


class PCDError(ValueError):
    pass


def _parse_header(f):
    hdr, n_lines = {}, 0
    while True:
        line = f.readline()
        if not line:
            raise PCDError("PCD header ends before a DATA line")
        n_lines += 1
        text = line.decode("ascii", "replace").strip()
        if not text or text.startswith("#"):
            continue
        key, _, rest = text.partition(" ")
        hdr[key.upper()] = rest.split()
        if key.upper() == "DATA":
            break
        if n_lines > 64:
            raise PCDError("PCD header too long")
    for k in ("FIELDS", "SIZE", "TYPE", "WIDTH", "HEIGHT", "DATA"):
        if k not in hdr:
            raise PCDError("PCD header lacks %s" % k)
    fields = hdr["FIELDS"]
    sizes = [int(v) for v in hdr["SIZE"]]
    types = hdr["TYPE"]
    counts = [int(v) for v in hdr.get("COUNT", ["1"] * len(fields))]
    if not (len(fields) == len(sizes) == len(types) == len(counts)):
        raise PCDError("FIELDS / SIZE / TYPE / COUNT disagree")
    width, height = int(hdr["WIDTH"][0]), int(hdr["HEIGHT"][0])
    points = int(hdr.get("POINTS", [width * height])[0])
    if points != width * height:
        raise PCDError("POINTS %d != WIDTH x HEIGHT %d" % (points, width * height))
    return fields, sizes, types, counts, width, height, hdr["DATA"][0].lower()


def _assemble(cols, n):
    out = np.zeros(n, dtype=POINT)
    for name in ("x", "y", "z"):
        if name not in cols:
            raise PCDError("PCD file has no field %r" % name)
        out[name] = cols[name].astype(np.float32)
    if "rgba" in cols:
        out["rgba"] = cols["rgba"].astype(np.uint32)
    elif "rgb" in cols:
        c = cols["rgb"]
        out["rgba"] = c.astype(np.float32).view(np.uint32) if c.dtype.kind == "f" else c.astype(np.uint32)
    return out


def loadPCDFile(path):
    """pcl::io::loadPCDFile<pcl::PointXYZRGBA>: returns (points, width, height)."""
    with open(path, "rb") as f:
        fields, sizes, types, counts, width, height, data = _parse_header(f)
        n = width * height
        if data == "ascii":
            cols = {}
            rows = [ln.split() for ln in f.read().decode("ascii", "replace").splitlines() if ln.strip()]
            if len(rows) < n:
                raise PCDError("PCD file holds %d of %d points" % (len(rows), n))
            col = 0
            for name, size, typ, cnt in zip(fields, sizes, types, counts):
                if cnt == 1 and name in ("x", "y", "z", "rgb", "rgba"):
                    vals = [r[col] for r in rows[:n]]
                    if name == "rgb" and typ == "F":
                        # PCL prints the packed colour as a float: its BITS are the colour
                        cols[name] = np.array([float(v) for v in vals], dtype=np.float32)
                    elif typ == "F":
                        cols[name] = np.array([float(v) for v in vals], dtype=np.float64)  # "nan" parses as NaN
                    else:
                        cols[name] = np.array([int(v) for v in vals], dtype=np.int64)
                col += cnt
            return _assemble(cols, n), width, height
        if data == "binary":
            dt = []
            for k, (name, size, typ, cnt) in enumerate(zip(fields, sizes, types, counts)):
                if (typ, size) not in _NP:
                    raise PCDError("unsupported field type %s%d" % (typ, size))
                dt.append(("f%d_%s" % (k, name), _NP[(typ, size)], (cnt,)) if cnt != 1 else ("f%d_%s" % (k, name), _NP[(typ, size)]))
            dt = np.dtype(dt)
            raw = f.read(n * dt.itemsize)
            if len(raw) < n * dt.itemsize:
                raise PCDError("PCD file holds %d of %d bytes of point data" % (len(raw), n * dt.itemsize))
            rec = np.frombuffer(raw, dtype=dt, count=n)
            cols = {}
            for k, (name, cnt) in enumerate(zip(fields, counts)):
                if cnt == 1 and name in ("x", "y", "z", "rgb", "rgba"):
                    cols[name] = rec["f%d_%s" % (k, name)]
            return _assemble(cols, n), width, height
        raise PCDError("DATA %s is not supported (ascii and binary are)" % data)


def savePCDFile(path, points, binary=False, width=None, height=1):
    """pcl::PCDWriter::write<pcl::PointXYZRGBA>(path, cloud, binary) (ref: src/create_model.cpp:223 passes false)."""
    pts = np.ascontiguousarray(points, dtype=POINT)
    n = len(pts)
    width = n if width is None else int(width)
    if width * int(height) != n:
        raise PCDError("width x height != number of points")
    header = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgba\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n"
              "WIDTH %d\nHEIGHT %d\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (width, int(height), n, "binary" if binary else "ascii"))
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        if binary:
            f.write(pts.tobytes())
        else:
            def fmt(v):
                return "nan" if np.isnan(v) else ("inf" if v == np.inf else ("-inf" if v == -np.inf else repr(float(np.float32(v)))))
            lines = ["%s %s %s %d" % (fmt(p["x"]), fmt(p["y"]), fmt(p["z"]), int(p["rgba"])) for p in pts]
            f.write(("\n".join(lines) + ("\n" if lines else "")).encode("ascii"))
