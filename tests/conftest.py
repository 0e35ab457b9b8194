import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and load libdagnn_sm100.so — nvcc cross-compiles without a GPU."""
    from dagnn_b200 import _lib
    _lib.build_library()
    return _lib.lib()
