import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def native_built():
    """Build the C-ABI library once per session if it is not there yet (nvcc cross-compiles without a GPU)."""
    from calamity_b200 import _native

    if not os.path.exists(_native.lib_path()):
        import __graft_entry__

        __graft_entry__.build()
    return _native.load()
