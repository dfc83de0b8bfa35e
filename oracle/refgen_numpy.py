"""CPU oracle (test infrastructure only) of the NMPC reference generation -- SURVEY.md section 8f-2.

numpy float64 restatement of ndp_nmpc/scripts/pt_pub/: piecewise-polynomial evaluation with the hover
branch after the end (base_pt_publisher.py:81-148), differential flatness (pt_publisher.py:188-248) and the
state/input packing traj_full_pt_2_x_u (pt_publisher.py:124-147).  quaternion_from_matrix restates the ROS
`tf` package's algorithm [EXT] (tf_conversions is not in /root/reference).  Pinned against
tests/golden/refgen_golden.npz (outputs of the reference's own functions, tests/golden/make_refgen_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

MASS, GRAVITY = 1.4844, 9.81  # params/fhnp_params.py:9,12


def _poly(c, k, s, t_seg):
    """k-th real-time derivative of sum_j c_j s^j at normalised time s (base_pt_publisher.py:138-146)."""
    n = len(c) - 1
    acc = 0.0
    for j in range(k, n + 1):
        acc += c[j] * math.perm(j, k) * s ** (j - k)
    return acc / t_seg**k


def quaternion_from_matrix(R):
    """(x, y, z, w) of a 3x3 rotation, ROS tf.transformations.quaternion_from_matrix [EXT]."""
    t = R[0, 0] + R[1, 1] + R[2, 2] + 1.0
    q = np.empty(4)
    if t > 1.0:
        q[3] = t; q[2] = R[1, 0] - R[0, 1]; q[1] = R[0, 2] - R[2, 0]; q[0] = R[2, 1] - R[1, 2]
    else:
        i, j, k = 0, 1, 2
        if R[1, 1] > R[0, 0]:
            i, j, k = 1, 2, 0
        if R[2, 2] > R[i, i]:
            i, j, k = 2, 0, 1
        t = R[i, i] - (R[j, j] + R[k, k]) + 1.0
        q[i] = t; q[j] = R[i, j] + R[j, i]; q[k] = R[k, i] + R[i, k]; q[3] = R[k, j] - R[j, k]
    return q * (0.5 / math.sqrt(t))


def flat_point(traj, t):
    """(pos, vel, acc, jerk, yaw, yaw_dot) at time t; hover at final_pt after the end."""
    tc = traj.t_cum
    if t >= tc[-1]:
        return traj.final_pt.copy(), np.zeros(3), np.zeros(3), np.zeros(3), 0.0, 0.0
    i = int(np.argwhere(tc > t)[0].item()) - 1
    ts = tc[i + 1] - tc[i]
    s = (t - tc[i]) / ts
    d = [np.array([_poly(c[i], k, s, ts) for c in (traj.cx, traj.cy, traj.cz)]) for k in range(4)]
    return d[0], d[1], d[2], d[3], _poly(traj.cyaw[i], 0, s, ts), _poly(traj.cyaw[i], 1, s, ts)


def full_state(pos, vel, acc, jerk, yaw, yaw_dot):
    """x[10] = (p, v, qw, qx, qy, qz), u[4] = (wx, wy, wz, c) with c = collective force / mass."""
    t_des = np.array([acc[0] + 0.0, acc[1] + 0.0, acc[2] + GRAVITY])
    tn = np.linalg.norm(t_des)
    z_b = t_des / tn
    u1 = tn * MASS
    x_c = np.array([math.cos(yaw), math.sin(yaw), 0.0])
    zx = np.cross(z_b, x_c)
    y_b = zx / np.linalg.norm(zx)
    x_b = np.cross(y_b, z_b)
    R = np.stack([x_b, y_b, z_b], 1)
    h_om = (MASS / u1) * (jerk - (z_b @ jerk) * z_b)
    p, q, r = -(h_om @ y_b), h_om @ x_b, yaw_dot * z_b[2]
    qx, qy, qz, qw = quaternion_from_matrix(R)
    return np.array([*pos, *vel, qw, qx, qy, qz]), np.array([p, q, r, u1 / MASS])


def ref_point(traj, t):
    return full_state(*flat_point(traj, t))


def horizon(traj, t0, N=20, th_pred=0.1, offset=None):
    """xr[N+1,10], ur[N,4]: the reference evaluated at t0 + k th_pred (SURVEY.md section D set-up)."""
    pts = [ref_point(traj, t0 + k * th_pred) for k in range(N + 1)]
    xr = np.array([p[0] for p in pts]); ur = np.array([p[1] for p in pts[:N]])
    if offset is not None:
        xr[:, 0:3] += np.asarray(offset)
    return xr, ur
