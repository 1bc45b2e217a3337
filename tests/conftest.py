import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Both shared libraries are built in-tree: the product (nvcc, sm_100a) and the checker (gcc)."""
    from oracle import binding as ob
    ob.build()
    import heros_b200
    if not os.path.exists(heros_b200._lib.LIB_PATH):
        heros_b200.build()
    yield
