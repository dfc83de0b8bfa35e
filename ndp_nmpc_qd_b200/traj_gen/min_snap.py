"""Closed-form minimum-snap / minimum-acceleration polynomial fit through way-points (host side, numpy).

Same linear system as the reference planner (cmd_pc/scripts/traj_gen/polym_optimizer.py:41-105 and
traj_generator.py:42-101): per segment a polynomial of order n = 2 Nd - 1 in NORMALISED time s in [0, 1]
(Nd = 4 snap for x/y/z, Nd = 2 acceleration for yaw), end-point interpolation, zero derivatives 1..Nd-1 at
both ends of the path, continuity of derivatives 1..n-1 (in normalised time) at interior way-points;
segment durations t = distance / mean speed.  Produces the TrajCoefficients message content
(ndp_nmpc/msg/TrajCoefficients.msg) consumed by the reference generator (`RefGen`).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def basis_row(n: int, k: int, s: float) -> np.ndarray:
    """d^k/ds^k of (1, s, ..., s^n)."""
    j = np.arange(n + 1, dtype=np.float64)
    fall = np.ones(n + 1)
    for d in range(k):
        fall *= np.maximum(j - d, 0.0)
    return fall * np.power(float(s), np.maximum(j - k, 0.0))


def fit(wpts: np.ndarray, nd: int) -> np.ndarray:
    """Coefficients [m, n+1] (ascending powers of s) of the m = len(wpts)-1 segments."""
    w = np.asarray(wpts, dtype=np.float64)
    m, n = len(w) - 1, 2 * nd - 1
    q = n + 1
    A, b = np.zeros((m * q, m * q)), np.zeros(m * q)
    r = 0
    for i in range(m):  # start points, then end points
        A[r, i * q:(i + 1) * q] = basis_row(n, 0, 0.0); b[r] = w[i]; r += 1
    for i in range(m):
        A[r, i * q:(i + 1) * q] = basis_row(n, 0, 1.0); b[r] = w[i + 1]; r += 1
    for k in range(1, nd):  # rest at the start of the path
        A[r, 0:q] = basis_row(n, k, 0.0); r += 1
    for k in range(1, nd):  # ... and at its end
        A[r, (m - 1) * q:m * q] = basis_row(n, k, 1.0); r += 1
    for i in range(m - 1):  # smooth junctions
        for k in range(1, n):
            A[r, i * q:(i + 1) * q] = basis_row(n, k, 1.0)
            A[r, (i + 1) * q:(i + 2) * q] = -basis_row(n, k, 0.0)
            r += 1
    return np.linalg.solve(A, b).reshape(m, q)


@dataclass
class Trajectory:
    t_cum: np.ndarray     # [m+1]
    cx: np.ndarray        # [m, 8]
    cy: np.ndarray
    cz: np.ndarray
    cyaw: np.ndarray      # [m, 4]
    final_pt: np.ndarray  # [3]

    @property
    def duration(self) -> float:
        return float(self.t_cum[-1])


def plan(pos: np.ndarray, yaw_deg: np.ndarray, vel: np.ndarray) -> Trajectory:
    pos = np.asarray(pos, dtype=np.float64)
    d = np.linalg.norm(pos[1:] - pos[:-1], axis=1)
    v = np.asarray(vel, dtype=np.float64)
    t_seg = d / ((v[:-1] + v[1:]) / 2)  # traj_generator.py:55-64
    return Trajectory(np.insert(np.cumsum(t_seg), 0, 0.0), fit(pos[:, 0], 4), fit(pos[:, 1], 4), fit(pos[:, 2], 4),
                      fit(np.radians(yaw_deg), 2), pos[-1].copy())


def plan_named(name: str) -> Trajectory:
    from .paths import PATHS

    p = PATHS[name]
    return plan(p["pos"], p["yaw_deg"], p["vel"])
