import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

REF_TESTS = "/root/reference/tests"           # only exists in the build container, never on the GPU box
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def have_gpu():
    try:
        import msamtools_b200 as m
        return m._lib.load().msg_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    orc.load()
    return orc


needs_reference = pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="/root/reference not present (GPU box)")
