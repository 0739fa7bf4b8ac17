"""The C-ABI library loads on a CPU-only box and exports every symbol include/pft/pft.h declares; the
ctypes binding covers all of them; compute entry points fail loudly without a device."""
import ctypes as C
import re

import pytest

from pcl_tracking_b200 import _capi as capi


def _declared():
    src = open(capi.HEADER_PATH).read()
    return sorted(set(re.findall(r"PFT_API\s+[\w\s\*]+?\b(pft_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 60
    lib = C.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libpft.so does not export %s" % n
        assert n in capi.SIGNATURES, "ctypes binding lacks %s" % n
    assert sorted(capi.SIGNATURES) == names


def test_library_has_no_oracle_or_torch_dependency():
    import subprocess
    out = subprocess.check_output(["ldd", capi.LIB_PATH]).decode()
    assert "pft_oracle" not in out and "torch" not in out and "nccl" not in out


def test_version_and_error_string():
    lib = capi.load()
    assert b"sm_100a" in lib.pft_version()
    assert isinstance(lib.pft_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    lib = capi.load()
    if lib.pft_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.pft_context_create(0, C.byref(h))
    assert rc == capi.ERR_CUDA
    assert b"no CPU fallback" in lib.pft_last_error()
    from pcl_tracking_b200 import pcl
    with pytest.raises(capi.PftError):
        pcl.Context(0)


def test_null_arguments_are_rejected():
    lib = capi.load()
    assert lib.pft_tracker_compute(None) == capi.ERR_INVALID
    assert lib.pft_tracker_set_i(None, 0, 0) == capi.ERR_INVALID
    assert lib.pft_cloud_create(None, None) == capi.ERR_INVALID
    assert lib.pft_compute_batch(None, -1) == capi.ERR_INVALID
