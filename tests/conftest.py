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
    """max over the batch of ||a-b||_inf / max(||b||_inf, 1) (SURVEY.md 8d parity gate)."""
    a = np.asarray(a, dtype=np.float64).reshape(a.shape[0], -1)
    b = np.asarray(b, dtype=np.float64).reshape(b.shape[0], -1)
    return float(np.max(np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1.0)))
