import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build libtmgpu.so and the oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    from tokenmonster_b200 import _lib

    if not os.path.exists(_lib.SO_PATH):
        g.build()
    from oracle import oracle

    oracle.build()
    return True
