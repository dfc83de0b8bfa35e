"""Hover-throttle hook kept on the boundary (host side, numpy, vectorised over quadrotors).

Reference: hv_throttle_est/hover_throttle_estimator.py:15-53 (2-state Kalman filter on
[f_collect, k_throttle]), differentiator.py:3-23 (Tustin differentiator, tau = 0.05),
alpha_filter.py:11-20.  The filter is a few scalar operations per quadrotor per 20 ms, so it stays
on the host; `batch` > 1 runs the same recursion for many quadrotors at once with the 2x2 algebra
written out in closed form (no per-element Python loop).
"""
from __future__ import annotations

import numpy as np

from ..params import estimator_params as EP


class AlphaFilter:
    """y[k] = alpha * y[k-1] + (1 - alpha) * u[k]   (alpha_filter.py:11-20)."""

    def __init__(self, alpha=0.5, y0=0.0):
        self.alpha = alpha
        self.y = y0

    def update(self, u):
        self.y = self.alpha * self.y + (1 - self.alpha) * u
        return self.y


class Differentiator:
    """Tustin-rule differentiator (differentiator.py:3-23)."""

    def __init__(self, Ts, batch: int | None = None):
        z = 0.0 if batch is None else np.zeros(batch)
        self.x_delay_1 = z
        self.x_dot_delay_1 = z
        tau = 0.05
        self.a1 = (2.0 * tau - Ts) / (2.0 * tau + Ts)
        self.a2 = 2.0 / (2.0 * tau + Ts)

    def update(self, x):
        x_dot = self.a1 * self.x_dot_delay_1 + self.a2 * (x - self.x_delay_1)
        # batched callers may hand in one preallocated buffer that they overwrite every tick: keep a private copy
        self.x_delay_1 = np.array(x, dtype=np.float64, copy=True) if isinstance(x, np.ndarray) else x
        self.x_dot_delay_1 = x_dot
        return x_dot


class HoverThrottleEstimator:
    """update(vz, throttle) -> (k_throttle, x, P)   (hover_throttle_estimator.py:37-53).

    Model: x = [f_collect, k_throttle], Phi = [[0, thr], [0, 1]], H = [1/m, 0], Q = diag(.1,.1),
    R = 1.225; the measurement is a_z + g with a_z from the differentiator.  The update only runs
    while 0.1 < throttle < 1.  batch=None reproduces the reference's scalar interface (x is a
    [2,1] array, P is [2,2]); batch=B keeps x as [B,2], P as [B,2,2] and returns k_throttle[B].
    """

    def __init__(self, ts: float, batch: int | None = None) -> None:
        self.batch = batch
        B = 1 if batch is None else batch
        self.vz_diff = Differentiator(ts, batch)
        self._x = np.tile(np.array([0.0, EP.k_throttle_init]), (B, 1))
        self._P = np.tile(np.eye(2), (B, 1, 1))
        self.K = None
        self.R = EP.R
        self.Q = EP.Q

    @property
    def x(self):
        return self._x[0].reshape(2, 1) if self.batch is None else self._x

    @property
    def P_mtx(self):
        return self._P[0] if self.batch is None else self._P

    def update(self, vz, throttle):
        az = self.vz_diff.update(vz)
        thr = np.atleast_1d(np.asarray(throttle, dtype=np.float64))
        z = np.atleast_1d(np.asarray(az, dtype=np.float64)) + EP.gravity
        on = (thr > 0.1) & (thr < 1.0)
        if np.any(on):
            x, P = self._x, self._P
            q0, q1 = self.Q[0, 0], self.Q[1, 1]
            p11 = P[:, 1, 1]
            # P^- = Phi P Phi' + Q with Phi = [[0, thr], [0, 1]]
            m00 = thr * thr * p11 + q0
            m01 = thr * p11
            m11 = p11 + q1
            im = 1.0 / EP.mass
            s = m00 * im * im + self.R  # H P^- H' + R
            k0, k1 = m00 * im / s, m01 * im / s
            xp0, xp1 = thr * x[:, 1], x[:, 1]  # x^- = Phi x
            innov = z - xp0 * im
            xn = np.stack([xp0 + k0 * innov, xp1 + k1 * innov], 1)
            # P = (I - K H) P^-
            Pn = np.empty_like(P)
            Pn[:, 0, 0] = (1 - k0 * im) * m00
            Pn[:, 0, 1] = (1 - k0 * im) * m01
            Pn[:, 1, 0] = m01 - k1 * im * m00
            Pn[:, 1, 1] = m11 - k1 * im * m01
            self._x = np.where(on[:, None], xn, x)
            self._P = np.where(on[:, None, None], Pn, P)
            self.K = np.stack([k0, k1], 1)
        k_throttle = self._x[:, 1]
        if self.batch is None:
            return float(k_throttle[0]), self.x, self.P_mtx
        return k_throttle.copy(), self.x, self.P_mtx


def nmpc_u_to_thrust(c, k_throttle, mass=EP.mass):
    """PX4 thrust command from the NMPC collective acceleration: c * m / k_throttle, 0 when the
    estimate is 0 (nmpc_node.py:273-283)."""
    c = np.asarray(c, dtype=np.float64)
    k = np.asarray(k_throttle, dtype=np.float64)
    return np.where(k != 0, c * mass / np.where(k != 0, k, 1.0), 0.0)
