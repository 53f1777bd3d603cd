import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built():
    """The in-tree native libraries (built by __graft_entry__.build(); rebuilt here if missing)."""
    lib = os.path.join(ROOT, "live-video-magnification_b200", "libmagcore_b200.so")
    hc = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
    if not (os.path.exists(lib) and os.path.exists(hc)):
        import __graft_entry__ as g
        g.build()
    return lib, hc
