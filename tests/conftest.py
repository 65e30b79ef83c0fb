import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


def rel_err(a, b):
    """Norm-relative error max|a-b| / max|b| (the parity metric of DESIGN.md)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.fixture(scope="session")
def dev():
    """cuda:0 on the B200 box.  The GPU tests fail (not skip) if the CUDA library is missing."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container (run with -m gpu on the B200 box)")
    from nrse_b200 import _lib
    _lib.load()
    _lib.check(_lib.load().nrse_check_device(), "nrse_check_device")
    return torch.device("cuda:0")
