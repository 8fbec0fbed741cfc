import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ofim():
    """CPU oracle (checker)."""
    from oracle import fimera

    return fimera


@pytest.fixture(scope="session")
def gfim():
    """CUDA library through the fimera-compatible C-ABI shim (the product)."""
    import chimera_b200.fimera as f

    if f.device_count() == 0:
        pytest.skip("no CUDA device")
    return f
