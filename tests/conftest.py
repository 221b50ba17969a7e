import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(ROOT / "tests" / "golden" / "hotpath_v1.npz")


@pytest.fixture(scope="session")
def lib_built():
    """Builds the C-ABI library in-tree if needed (nvcc cross-compiles without a GPU)."""
    from hifimeth_b200 import build

    return build.build()
