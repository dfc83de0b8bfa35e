import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Compile libndp_nmpc_b200.so if it is missing or stale (nvcc cross-compiles without a GPU)."""
    from ndp_nmpc_qd_b200 import build

    return build.build()


@pytest.fixture(scope="session")
def c_oracle():
    from oracle.c_oracle import COracle

    return COracle()


@pytest.fixture(scope="session")
def mlp_weights():
    from oracle import mlp_numpy

    from ndp_nmpc_qd_b200.dnwash_nn_est.downwash_nn import DEFAULT_WEIGHTS

    return mlp_numpy.load_npz(DEFAULT_WEIGHTS)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_err(a, b):
    """Parity gate, PER COMPONENT: max over every element of |a - b| / max(|b|, 1) -- so a body rate is held to 1e-4
    rad/s even though the collective acceleration next to it is ~10 (north_star: u0 and predicted trajectories within
    1e-4 relative in fp32)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))) if a.size else 0.0
