import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "heimdall-vision_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session", params=["fused", "global"])
def detector(request):
    """One CUDA context per CCL path for the whole GPU session; fails loudly (no skip) if the extension or the GPU is
    missing.  "fused" = default configuration (per-frame shared-memory CCL kernel with automatic fallback to the
    global-memory kernels for dense frames), "global" = global-memory CCL kernels only."""
    import heimdall_core
    det = heimdall_core.Detector(0, max_defects_per_frame=8192, global_ccl=(request.param == "global"))
    yield det
    det.close()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
