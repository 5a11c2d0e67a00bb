import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# Fresh device buffers of the library start as 0x7F.. instead of whatever the allocator hands out (read once, at the first
# allocation): a kernel that reads something it never wrote fails the same way on every box, not only behind an unlucky test
os.environ.setdefault("SZ_DEBUG_POISON", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """the product library and the checker are built once per session (no-ops when up to date)"""
    import __graft_entry__ as g
    g.build()
    yield
