import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def ctx():
    """A libgingr_cuda context on cuda:0.  Fails loudly (no fallback) when the library or GPU is missing."""
    from gingr_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(float(np.max(np.abs(b))), floor, 1e-300)
    return float(np.max(np.abs(a - b)) / scale)
