"""Synthetic reference horizons for tests and benchmarks (host side, numpy).

The reference obtains (xr[21,10], ur[20,4]) from a min-snap polynomial through
the way-points of cmd_pc/path_config/eight_*.yaml followed by the differential
flatness map of ndp_nmpc/scripts/pt_pub/pt_publisher.py:188-248.  Here the path
is an analytic figure-eight (lemniscate) with the same envelope as those YAMLs
(centre (1,1,5), 10 m x 5 m x 3 m half-extents; "high_dyn" ~ 9.6 m/s peak,
"low" ~ 1 m/s peak), pushed through the same flatness map, so that the NMPC
sees horizons of the shape and dynamic range the reference's node produces.
"""
from __future__ import annotations

import numpy as np

GRAVITY = 9.81  # params/fhnp_params.py:12
MASS = 1.4844  # params/fhnp_params.py:9

# name -> (period [s], half extents [m]); peak speed ~ 2*pi/period * |(10, 2*5, 3)|
PATHS = {
    "eight_high_dyn": (11.0, (10.0, 5.0, 3.0)),
    "eight_low": (110.0, (10.0, 5.0, 3.0)),
}


def figure_eight(t, name="eight_high_dyn", center=(1.0, 1.0, 5.0)):
    """Position and its first three time derivatives, each [..., 3]."""
    period, (ax, ay, az) = PATHS[name]
    w = 2.0 * np.pi / period
    t = np.asarray(t, dtype=np.float64)
    s, c = np.sin(w * t), np.cos(w * t)
    s2, c2 = np.sin(2 * w * t), np.cos(2 * w * t)
    pos = np.stack([center[0] + ax * s, center[1] + ay * s2, center[2] + az * s], -1)
    vel = np.stack([ax * w * c, 2 * ay * w * c2, az * w * c], -1)
    acc = np.stack([-ax * w * w * s, -4 * ay * w * w * s2, -az * w * w * s], -1)
    jerk = np.stack([-ax * w**3 * c, -8 * ay * w**3 * c2, -az * w**3 * c], -1)
    return pos, vel, acc, jerk


def quat_from_rot(R):
    """Rotation matrices [...,3,3] -> quaternions (w,x,y,z) with w >= 0
    (ROS convention used at pt_publisher.py:234-240)."""
    m00, m11, m22 = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    w = 0.5 * np.sqrt(np.maximum(1.0 + m00 + m11 + m22, 1e-12))
    x = (R[..., 2, 1] - R[..., 1, 2]) / (4 * w)
    y = (R[..., 0, 2] - R[..., 2, 0]) / (4 * w)
    z = (R[..., 1, 0] - R[..., 0, 1]) / (4 * w)
    q = np.stack([w, x, y, z], -1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def diff_flatness(pos, vel, acc, jerk, yaw=0.0, yaw_dot=0.0):
    """Full state/input from flat outputs (restates pt_publisher.py:188-248).

    Returns x[...,10] = (p, v, qw, qx, qy, qz) and u[...,4] = (wx, wy, wz, c) with c the
    collective acceleration (force / mass, pt_publisher.py:140-147).
    """
    t_des = acc + np.array([0.0, 0.0, GRAVITY])
    c = np.linalg.norm(t_des, axis=-1, keepdims=True)
    z_b = t_des / c
    x_c = np.broadcast_to(np.array([np.cos(yaw), np.sin(yaw), 0.0]), z_b.shape)
    y_b = np.cross(z_b, x_c)
    y_b = y_b / np.linalg.norm(y_b, axis=-1, keepdims=True)
    x_b = np.cross(y_b, z_b)
    R = np.stack([x_b, y_b, z_b], -1)
    h_w = (jerk - np.sum(z_b * jerk, -1, keepdims=True) * z_b) / c
    p = -np.sum(h_w * y_b, -1)
    q = np.sum(h_w * x_b, -1)
    r = yaw_dot * z_b[..., 2]
    x = np.concatenate([pos, vel, quat_from_rot(R)], -1)
    u = np.stack([p, q, r, c[..., 0]], -1)
    return x, u


def reference_horizon(t0, N=20, th_pred=0.1, name="eight_high_dyn", center=(1.0, 1.0, 5.0)):
    """xr[B,N+1,10], ur[B,N,4] starting at phases t0[B] (reference nodes every th_pred)."""
    t0 = np.atleast_1d(np.asarray(t0, dtype=np.float64))
    t = t0[:, None] + th_pred * np.arange(N + 1)[None, :]
    x, u = diff_flatness(*figure_eight(t, name, center))
    return x, u[:, :N, :].copy()


def random_rotation_quat(rng, n, max_angle_rad):
    axis = rng.normal(size=(n, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    ang = rng.uniform(0.0, max_angle_rad, size=(n, 1))
    return np.concatenate([np.cos(ang / 2), np.sin(ang / 2) * axis], 1)


def quat_mul(a, b):
    aw, ax, ay, az = (a[..., i] for i in range(4))
    bw, bx, by, bz = (b[..., i] for i in range(4))
    return np.stack(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ],
        -1,
    )


def independent_problems(B, N=20, seed=0, scale=1.0, name=None, with_neighbour=False):
    """Config 3 of BASELINE.json: B independent single-quad problems.

    Trajectory phase t0 ~ U(0, period) on eight_high_dyn (50 %) / eight_low (50 %);
    x0 = xr_0 + perturbation (sigma_p = 0.1 m, sigma_v = 0.2 m/s, attitude rotated by
    U(0, 10 deg)) times `scale` (scale = 5 is the stress variant that activates bounds).
    with_neighbour adds a neighbour reference horizon 0.5-1.5 m above each ego, inside
    the 1 m horizontal gate, for the downwash MLP.
    Returns dict of float64 arrays: x0, xr, ur (+ other[B,N+1,10]).
    """
    rng = np.random.default_rng(seed)
    xr = np.zeros((B, N + 1, 10))
    ur = np.zeros((B, N, 4))
    which = rng.random(B) < 0.5 if name is None else np.full(B, name == "eight_high_dyn")
    t0_all = np.zeros(B)
    for nm, mask in (("eight_high_dyn", which), ("eight_low", ~which)):
        if mask.any():
            t0 = rng.uniform(0.0, PATHS[nm][0], size=int(mask.sum()))
            t0_all[mask] = t0
            xr[mask], ur[mask] = reference_horizon(t0, N, 0.1, nm)
    x0 = xr[:, 0, :].copy()
    x0[:, 0:3] += scale * 0.1 * rng.normal(size=(B, 3))
    x0[:, 3:6] += scale * 0.2 * rng.normal(size=(B, 3))
    dq = random_rotation_quat(rng, B, np.deg2rad(min(10.0 * scale, 170.0)))
    q = quat_mul(x0[:, 6:10], dq)
    x0[:, 6:10] = q / np.linalg.norm(q, axis=1, keepdims=True)
    out = dict(x0=x0, xr=xr, ur=ur, t0=t0_all, high_dyn=which)
    if with_neighbour:
        other = xr.copy()
        off = np.concatenate([rng.uniform(-0.5, 0.5, size=(B, 2)), rng.uniform(0.5, 1.5, size=(B, 1))], 1)
        other[:, :, 0:3] += off[:, None, :]
        other[:, :, 3:6] += 0.1 * rng.normal(size=(B, 1, 3))
        out["other"] = other
        out["other_offset"] = np.concatenate([off, other[:, 0, 3:6] - xr[:, 0, 3:6]], 1)  # constant (p, v) offset of the neighbour
    return out


def sliding_lists(w, n_ticks, ts=0.02, length=101):
    """The 50 Hz lists NMPCRefPublisher keeps (pt_pub/pt_publisher.py:57-103) for the problems of `independent_problems`:
    point i = the reference at t0 + ts * i, i < length + n_ticks; tick j sees points j .. j + length - 1, i.e. every 5th of
    them is the horizon at t0 + ts * j.  Returns float64 x_list [B, length + n_ticks, 10], u_list [.., 4] and, when the
    workload has a neighbour, other_list [.., 6] (its position / velocity columns)."""
    B = w["x0"].shape[0]
    n = length + n_ticks
    xl, ul = np.zeros((B, n, 10)), np.zeros((B, n, 4))
    for nm, mask in (("eight_high_dyn", w["high_dyn"]), ("eight_low", ~w["high_dyn"])):
        if mask.any():
            t = w["t0"][mask][:, None] + ts * np.arange(n)[None, :]
            xl[mask], ul[mask] = diff_flatness(*figure_eight(t, nm))
    out = dict(x_list=xl, u_list=ul)
    if "other_offset" in w:
        out["other_list"] = xl[:, :, 0:6] + w["other_offset"][:, None, :]
    return out


V_BOX = (4.5, 4.5, 1.4)  # ~79 % of the peak axis speeds of "eight_high_dyn" (5.7, 5.7, 1.7 m/s)


def velocity_box_problems(B, N=20, seed=0, v_max=V_BOX, sigma_p=0.05, sigma_v=0.05):
    """Feasible problems whose solution rides a VELOCITY bound (nmpc_body_rate_ctl.py:59-61,66: lbx/ubx on idx 3,4,5,
    stages 1..N-1), which the reference's own limits (+-20 m/s) never activate: trajectory phases of "eight_high_dyn"
    where the start speed is inside 85 % of the tightened box `v_max` on every axis while the reference exceeds 110 % of
    it later in the horizon, x0 = xr_0 + a small perturbation.  Use with v_min = -v_max, v_max = v_max.
    Returns dict of float64 arrays x0, xr, ur."""
    rng = np.random.default_rng(seed)
    vm = np.asarray(v_max, dtype=np.float64)
    xs, us = [], []
    have = 0
    while have < B:
        t0 = rng.uniform(0.0, PATHS["eight_high_dyn"][0], size=max(4 * B, 64))
        xr, ur = reference_horizon(t0, N, 0.1, "eight_high_dyn")
        sel = np.all(np.abs(xr[:, 0, 3:6]) <= 0.85 * vm, 1) & np.any(np.abs(xr[:, :, 3:6]).max(1) > 1.1 * vm, 1)
        xs.append(xr[sel]); us.append(ur[sel])
        have += int(sel.sum())
    xr, ur = np.concatenate(xs)[:B], np.concatenate(us)[:B]
    x0 = xr[:, 0, :].copy()
    x0[:, 0:3] += sigma_p * rng.normal(size=(B, 3))
    x0[:, 3:6] += sigma_v * rng.normal(size=(B, 3))
    return dict(x0=x0, xr=xr, ur=ur)
