import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (restatement + unmodified reference when present)."""
    from oracle import oracle as O
    O.port()  # builds liboracle.so on first use
    return O


@pytest.fixture(scope="session")
def lz4_cases():
    return dict(np.load(os.path.join(GOLDEN, "lz4_cases.npz")))


@pytest.fixture(scope="session")
def zstd_cases():
    return dict(np.load(os.path.join(GOLDEN, "zstd_cases.npz")))


@pytest.fixture(scope="session")
def gpu_ctx():
    """The product: CUDA library through its C-ABI.  No fallback — missing library or GPU is a failure."""
    import zpack_b200
    ctx = zpack_b200.Context(0)
    yield ctx
    ctx.close()
