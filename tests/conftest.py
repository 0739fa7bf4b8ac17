import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Compile the C-ABI library (nvcc cross-compiles without a GPU) and the CPU oracle once per session."""
    from pcl_tracking_b200 import build as pft_build
    import oracle
    pft_build.build()
    oracle.build()
    yield
