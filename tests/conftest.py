import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pathintegral-qmc_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {
        "inst": np.load(os.path.join(d, "instances.npz")),
        "vec": np.load(os.path.join(d, "ref_vectors.npz")),
        "dist": np.load(os.path.join(d, "ref_distributions.npz")),
    }


@pytest.fixture(scope="session")
def dev():
    """The process-wide Device on cuda:0 (gpu tests only)."""
    from piqmc import device
    return device.default_device(0)
