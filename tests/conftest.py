import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """CPU restatement of the reference (oracle/liboracle.so) — the checker, never the product."""
    import pyoracle
    return pyoracle.load("port")


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled from /root/reference (oracle/_ref); skipped where it cannot exist."""
    import pyoracle
    if not pyoracle.available("reference") and not os.path.isdir(pyoracle.REFERENCE_ROOT):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return pyoracle.load("reference")
