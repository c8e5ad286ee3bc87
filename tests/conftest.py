import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    from lrzip_next_b200 import Context
    from lrzip_next_b200.api import LrzGpuError
    try:
        c = Context(0)
    except LrzGpuError as e:
        if e.code == -4:  # LRZGPU_ENODEV: the product has no CPU fallback, so GPU tests cannot run on this box
            pytest.skip("no CUDA device")
        raise
    yield c
    c.close()
