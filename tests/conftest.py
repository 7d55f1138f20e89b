import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The tests exercise the real shared library (host I/O, ROI selector, ABI symbols on CPU; kernels on the GPU
    box), so make sure it is built and current; nvcc cross-compiles without a GPU.  The package itself never builds
    or falls back silently: without the library every call raises."""
    from epilogos_b200 import build
    build.build()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(GOLDEN / (name + ".npz"), allow_pickle=False)
    return load
