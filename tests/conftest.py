import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); runs through libigl_b200/libiglb200.so")


@pytest.fixture(scope="session")
def gpu_api():
    """The PRODUCT library on a real GPU.  Fails (does not skip) when the extension is missing."""
    import libigl_b200
    api = libigl_b200.lib()
    assert api.device_count() > 0, "no CUDA device visible"
    assert "EMULATION" not in api.version()
    return api


@pytest.fixture(scope="session")
def emul_api():
    """Host emulation of the kernels (tests/emul.py) — logic checks without a GPU."""
    import emul
    return emul.api()
