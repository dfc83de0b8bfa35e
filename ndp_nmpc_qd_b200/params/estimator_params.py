"""Hover-throttle estimator constants (reference: params/estimator_params.py:13-18)."""
import numpy as np

from .fhnp_params import gravity, mass  # noqa: F401

k_throttle_init = 50.0
ts_est = 0.02
R = 1.225
Q = np.diag([0.1, 0.1])
