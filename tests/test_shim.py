"""The C++ PCL-named shim (include/pft/pcl_shim.hpp) and the offline driver compile with plain g++ against
libpft.so; without a GPU the driver fails loudly (exit 3, "no CPU fallback")."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "auto_tracking_offline")


def build_driver():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s", "CXX=/usr/bin/g++"])
    return EXE


def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = build_driver()
    from pcl_tracking_b200 import _capi
    if _capi.load().pft_device_count() > 0:
        pytest.skip("a CUDA device is present")
    f = tmp_path / "x.raw"
    f.write_bytes(b"\0" * 64)
    r = subprocess.run([exe, str(f), "1", str(f)], capture_output=True, text=True)
    assert r.returncode == 3
    assert "no CPU fallback" in r.stderr
