import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """The oracle (gcc) is rebuilt if stale; the CUDA library must already be built in-tree."""
    import oracle
    oracle.build()
    from symbanafis_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    yield
