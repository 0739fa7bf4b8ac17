"""PCD reader / writer (the data format either side of the path: model clouds written by the model builder,
ref: src/create_model.cpp:209-230, and frames replayed offline).  Host-side only: runs without a GPU."""
import struct

import numpy as np
import pytest

import oracle
from pcl_tracking_b200 import pcd


def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    if n > 4:
        xyz[1, 0] = np.nan
        xyz[3] = (np.inf, -np.inf, 1e-30)
    return oracle.make_points(xyz, rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32))


@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("n", [0, 1, 257])
def test_pcd_round_trip_is_bit_exact(tmp_path, binary, n):
    pts = _cloud(n, 3 + n)
    path = tmp_path / "c.pcd"
    pcd.savePCDFile(path, pts, binary=binary)
    got, w, h = pcd.loadPCDFile(path)
    assert (w, h) == (n, 1)
    assert got.tobytes() == pts.tobytes()


def test_pcd_organised_cloud_and_foreign_fields(tmp_path):
    # an organised 3 x 2 cloud with a normal field between xyz and a float-packed rgb, as other PCL tools write it
    path = tmp_path / "o.pcd"
    rows = []
    want = np.zeros(6, dtype=pcd.POINT)
    for i in range(6):
        rgba = 0x00102030 + i
        as_float = struct.unpack("<f", struct.pack("<I", rgba))[0]
        want[i] = (0.5 * i, -1.0 * i, 2.0, rgba)
        rows.append("%r %r %r 0 0 1 %s" % (0.5 * i, -1.0 * i, 2.0, np.format_float_scientific(np.float32(as_float), unique=True)))
    path.write_text("# .PCD v0.7\nVERSION 0.7\nFIELDS x y z normal rgb\nSIZE 4 4 4 4 4\nTYPE F F F F F\nCOUNT 1 1 1 3 1\nWIDTH 3\nHEIGHT 2\n"
                    "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS 6\nDATA ascii\n" + "\n".join(rows) + "\n")
    got, w, h = pcd.loadPCDFile(path)
    assert (w, h) == (3, 2)
    assert got.tobytes() == want.tobytes()
    # the same records in binary
    rec = np.zeros(6, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("n", "<f4", (3,)), ("rgb", "<u4")])
    rec["x"], rec["y"], rec["z"], rec["rgb"] = want["x"], want["y"], want["z"], want["rgba"]
    hdr = path.read_text().split("DATA")[0] + "DATA binary\n"
    path.write_bytes(hdr.encode() + rec.tobytes())
    got, w, h = pcd.loadPCDFile(path)
    assert got.tobytes() == want.tobytes()


def test_pcd_errors(tmp_path):
    p = tmp_path / "bad.pcd"
    p.write_text("VERSION 0.7\nFIELDS x y\nSIZE 4 4\nTYPE F F\nCOUNT 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2\n")
    with pytest.raises(pcd.PCDError):
        pcd.loadPCDFile(p)                      # no z field
    p.write_text("VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary_compressed\n")
    with pytest.raises(pcd.PCDError):
        pcd.loadPCDFile(p)
    p.write_text("VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA ascii\n1 2 3\n")
    with pytest.raises(pcd.PCDError):
        pcd.loadPCDFile(p)                      # truncated


def test_cpp_shim_pcd_io_interoperates(tmp_path):
    """pcl::io::loadPCDFile / pcl::PCDWriter::write / savePCDFileBinary of the C++ shim read and write the same files
    bit for bit (host-side only: compiles and runs without a GPU)."""
    import os
    import subprocess
    from pcl_tracking_b200 import build as pft_build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.dirname(pft_build.build())
    src = tmp_path / "pcd_check.cpp"
    src.write_text(r'''
#include <pft/pcl_shim.hpp>
int main(int argc, char** argv) {
  typedef pcl::PointXYZRGBA P;
  pcl::PointCloud<P> in, a, b;
  if (argc < 4 || pcl::io::loadPCDFile(argv[1], in) != 0) return 1;
  pcl::PCDWriter w;
  if (w.write(argv[2], in, false) != 0) return 2;                 // ref: src/create_model.cpp:223
  if (pcl::io::savePCDFileBinary(argv[3], in) != 0) return 3;
  if (pcl::io::loadPCDFile(argv[2], a) != 0 || pcl::io::loadPCDFile(argv[3], b) != 0) return 4;
  return (a.points.size() == in.points.size() && b.points.size() == in.points.size()) ? 0 : 5;
}
''')
    exe = tmp_path / "pcd_check"
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), "-L", lib_dir, "-lpft",
                           "-Wl,-rpath," + lib_dir])
    pts = _cloud(300, 8)
    pcd.savePCDFile(tmp_path / "in.pcd", pts, binary=True)
    subprocess.check_call([str(exe), str(tmp_path / "in.pcd"), str(tmp_path / "a.pcd"), str(tmp_path / "b.pcd")])
    for name in ("a.pcd", "b.pcd"):
        got, w, h = pcd.loadPCDFile(tmp_path / name)
        assert got.tobytes() == pts.tobytes(), name
