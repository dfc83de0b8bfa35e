"""ORACLE (test infrastructure only): numpy restatement of the downwash MLP path.

Reference: dnwash_nn_est/nn_net.py:7-18 (Linear 6-128, ReLU, Linear 128-64, ReLU, Linear 64-128,
ReLU, Linear 128-3), dnwash_nn_est/downwash_nn.py:21-29 (features = (other - ego)[:, 0:6] computed
in float64 then cast to float32), gate ndp_nmpc_leader_node.py:65-76.  Pinned against outputs of
the reference's own torch module with the shipped weights (tests/golden/mlp_golden.npz, generated
by tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

KEYS = ("0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias", "6.weight", "6.bias")


def load_npz(path):
    raw = np.load(path)
    return {k: raw[k] for k in KEYS}


def mlp_forward(w: dict, x: np.ndarray, dtype=np.float32) -> np.ndarray:
    """x [M,6] -> [M,3]; arithmetic in `dtype` (float32 mirrors torch fp32 up to summation order)."""
    h = np.asarray(x, dtype=dtype)
    for i, k in enumerate(("0", "2", "4", "6")):
        h = h @ w[f"{k}.weight"].astype(dtype).T + w[f"{k}.bias"].astype(dtype)
        if i < 3:
            h = np.maximum(h, 0)
    return h


def downwash_update(w: dict, other_pred_x: np.ndarray, ego_pred_x: np.ndarray, dtype=np.float32) -> np.ndarray:
    """DownwashNN.update (downwash_nn.py:21-29): [n,10] x2 -> float32 [n,3]."""
    feat = (np.asarray(other_pred_x) - np.asarray(ego_pred_x))[:, 0:6].astype(np.float32)
    return mlp_forward(w, feat, dtype).astype(np.float32)


def gated_pairs(w: dict, ego, other, gate_xy=None, r_horiz=1.0, dtype=np.float64):
    """ego/other [P,n,10]; gate on the neighbour's node-0 horizontal distance to gate_xy [P,2]
    (ndp_nmpc_leader_node.py:65-68); zeros when the gate is closed (:75-76)."""
    P, n = ego.shape[:2]
    feat = (other - ego)[:, :, 0:6].astype(np.float32).reshape(P * n, 6)
    f = mlp_forward(w, feat, dtype).reshape(P, n, 3)
    if gate_xy is not None:
        d = other[:, 0, 0:2] - gate_xy
        on = (d[:, 0] ** 2 + d[:, 1] ** 2) < r_horiz**2
        f = f * on[:, None, None]
    return f


def swarm_forces(w: dict, traj, ego_begin, n_ego, odom_xy=None, r_horiz=1.0, dtype=np.float64):
    """traj [n_all,n,6] float32: f_i = sum_{j != i, gated} MLP(traj_j - traj_i)."""
    n_all, n = traj.shape[:2]
    out = np.zeros((n_ego, n, 3), dtype=np.float64)
    r2 = np.float32(r_horiz * r_horiz)
    for i in range(n_ego):
        gi = ego_begin + i
        exy = traj[gi, 0, 0:2] if odom_xy is None else odom_xy[i]
        d = traj[:, 0, 0:2] - exy
        on = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) < r2
        on[gi] = False
        js = np.nonzero(on)[0]
        if len(js):
            feat = (traj[js] - traj[gi][None]).reshape(-1, 6).astype(np.float32)
            out[i] = mlp_forward(w, feat, dtype).reshape(len(js), n, 3).sum(0)
    return out
