"""The three-quadrotor formation of BASELINE.json config 2 (`three_qd_ndp_nmpc`) without ROS.

The reference wires this scenario out of three node processes and the simulator
(ndp_nmpc/launch/three_qd_ndp_nmpc.launch:1-20):
  fhnp        NDPLeaderNode   NDP-NMPC controller + DownwashNN w.r.t. xiao_feng   (ndp_nmpc_leader_node.py:28-76)
  xiao_feng   FollowerNode    plain NMPC on leader reference + formation offset    (nmpc_follower_node.py:26-74)
  smile_boy   FollowerNode    same
  dop_sim     DopQdNode       MulQuadrotors plant, 3 agents, downwash / motor / battery on (three_qd_config.yaml)
This module restates the per-callback logic of those nodes as plain functions / small classes (each citing the
callback it follows; pinned to the reference's own methods by tests/golden/wire_golden.npz) and a deterministic
scheduler `ThreeQuadFormation` that replaces the ROS timers: plant 100 Hz, control 50 Hz, formation reference 20 Hz.
The controllers, the downwash observer and the plant are injected, so the same scenario runs on the drop-in CUDA
classes (product) and on the CPU oracles (tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from .hv_throttle_est import AlphaFilter
from .params import downwash_params as DP
from .params import estimator_params as EP
from .params import nmpc_params as CP

NAMES = ("fhnp", "xiao_feng", "smile_boy")


def leader_formation_refs(leader_odom_x: float):
    """NDPLeaderNode.pub_formation_ref_callback (ndp_nmpc_leader_node.py:49-58): formation offsets published at 20 Hz for
    (xiao_feng, smile_boy); xiao_feng moves from beside the leader to 0.5 m above it while |x - 1| > 2."""
    if abs(leader_odom_x - 1) > 2:
        return (0.0, 0.0, 0.5), (0.0, -1.0, 0.0)
    return (0.0, 1.0, 0.0), (0.0, -1.0, 0.0)


class FormationOffsetFilter:
    """FollowerNode.sub_formation_ref_callback (nmpc_follower_node.py:44-55): one AlphaFilter(alpha = 0.8) per axis,
    initialised with the first message, which is then also pushed through the filter."""

    def __init__(self, alpha: float = 0.8):
        self.alpha = alpha
        self._f: Optional[List[AlphaFilter]] = None
        self.value = np.array([1.0, 1.0, 0.5])  # FollowerNode.__init__ default (nmpc_follower_node.py:32)

    def update(self, msg_xyz: Sequence[float]) -> np.ndarray:
        if self._f is None:
            self._f = [AlphaFilter(alpha=self.alpha, y0=float(v)) for v in msg_xyz]
        self.value = np.array([f.update(float(v)) for f, v in zip(self._f, msg_xyz)])
        return self.value


def follower_reference(leader_xr: np.ndarray, leader_ur: np.ndarray, offset: Sequence[float]):
    """FollowerNode.sub_pred_callback (nmpc_follower_node.py:57-74): the leader's PredXU with the filtered formation
    offset added to the position of every node."""
    xr = np.array(leader_xr, dtype=np.float64, copy=True)
    xr[:, 0:3] += np.asarray(offset, dtype=np.float64)
    return xr, np.array(leader_ur, dtype=np.float64, copy=True)


def gate_open(other_x0_xy: Sequence[float], ego_odom_xy: Sequence[float], r_horiz: float = DP.r_horiz) -> bool:
    """ndp_nmpc_leader_node.py:65-68: the neighbour's node 0 strictly inside r_horiz of the ego's odometry position."""
    return (other_x0_xy[0] - ego_odom_xy[0]) ** 2 + (other_x0_xy[1] - ego_odom_xy[1]) ** 2 < r_horiz**2


def leader_disturb_force(downwash_update: Callable, other_xr: np.ndarray, ego_odom_xy: Sequence[float], ego_xr: np.ndarray) -> np.ndarray:
    """NDPLeaderNode.sub_xf_pred_callback (ndp_nmpc_leader_node.py:60-76): DownwashNN.update(other, ego reference) when
    the gate is open, zeros [n, 3] otherwise."""
    if gate_open(other_xr[0, 0:2], ego_odom_xy):
        return downwash_update(np.asarray(other_xr, dtype=np.float64), ego_xr)
    return np.zeros([other_xr.shape[0], 3])


def odom_to_x0(state_row: np.ndarray) -> np.ndarray:
    """dop_qd_node.py:131-148 + pt_publisher.py:106-122: plant state [35] -> x0 = (p, v, qw, qx, qy, qz)."""
    s = state_row
    return np.array([s[3], s[4], s[5], s[13], s[14], s[15], s[9], s[10], s[11], s[12]], dtype=np.float64)


def u0_to_cmd(u0: np.ndarray, k_throttle: float) -> np.ndarray:
    """nmpc_node.py:273-283 + dop_qd_node.py:162-166: (body rates, thrust = c m / k_throttle, 0 when k_throttle is 0)."""
    return np.array([u0[0], u0[1], u0[2], u0[3] * CP.mass / k_throttle if k_throttle != 0 else 0.0])


class ThreeQuadFormation:
    """Deterministic replay of the three-node formation against one plant.

    leader / followers: controller objects with the reference's interface (`reset(xr, ur)`,
    `update(x0, xr, ur[, f]) -> u0`); downwash_update: `DownwashNN.update`; plant_forward(ts_sim, state[3,35], cmd[3,4])
    -> new state (MulQuadrotors.forward semantics, dop_sim/scripts/quadrotor/mul_quadrotors.py:40-50);
    leader_reference(t) -> (xr[21,10], ur[20,4]) of the tracked trajectory at trajectory time t (get_nmpc_pts,
    pt_publisher.py:78-97).  make_estimator() -> an object with HoverThrottleEstimator's update(vz, thrust).

    Time base 0.01 s (ts_sim, three_qd_config.yaml:1).  Per sim step i: [i % 5 == 0] formation references (20 Hz);
    [i % 2 == 0] control tick (50 Hz): leader, then the followers on the PredXU the leader just published, then the
    leader's downwash update from xiao_feng's PredXU (used by the leader's NEXT tick -- in ROS the subscriber runs
    between two timer ticks); then one plant step.  Before `track()` the nodes hover on gen_fix_pt_ref
    (pt_publisher.py:40-55) with the hover-throttle filters running (nmpc_node.py:98-101); during tracking the
    filters are frozen (nmpc_node.py:146).
    """

    INIT_POS = np.array([[1.0, 1.0, 0.5], [1.0, 2.0, 0.5], [1.0, 0.0, 0.5]])  # three_qd_config.yaml:13-27

    START_TICK = (0, 3, 7)  # control tick at which each node's timers start (see below)

    def __init__(self, leader, followers, downwash_update: Callable, plant_forward: Callable, leader_reference: Callable,
                 make_estimator: Callable, ts_sim: float = 0.01, record: bool = False):
        """roslaunch starts the three node processes one after the other, so their controllers take over from the
        simulator's initial command (throttle 0.283, dop_qd_node.py:190) at different times: START_TICK models that.
        It is not cosmetic: with bit-identical quadrotors the simulator's pairwise downwash term
        (qd_dynamics.py:161-198, ~ (r_p / 4 dz)^2 for 0 < dz) is evaluated at dz ~ 1e-17 and overflows."""
        self.ctl = [leader, followers[0], followers[1]]
        self.downwash_update, self.plant_forward, self.leader_reference = downwash_update, plant_forward, leader_reference
        self.ts_sim = ts_sim
        self.est = [make_estimator() for _ in range(3)]
        self.k_throttle = [EP.k_throttle_init] * 3
        self.filters = [FormationOffsetFilter(), FormationOffsetFilter()]
        self.state = np.zeros((3, 35))
        self.state[:, 9] = -1.0   # ew = -1 is what the simulator starts with (dop_qd_node.py:174)
        self.state[:, 31:35] = 8.0  # kRPM (dop_qd_node.py:175)
        self.state[:, 3:6] = self.INIT_POS
        self.cmd = np.zeros((3, 4))
        self.cmd[:, 3] = 0.283  # dop_qd_node.py:190
        self.thrust = [0.0, 0.0, 0.0]  # AttitudeTarget().thrust before the first control tick
        self.xr = [None] * 3
        self.ur = [None] * 3
        self.disturb_force = np.zeros([CP.N_node + 1, 3])  # nmpc_node.py:52
        self.i = 0
        self.t_traj: Optional[float] = None
        self.log = [] if record else None
        self.dw_log = [] if record else None  # (other, ego, f) of every DownwashNN.update call that passed the gate
        for q in range(3):  # gen_fix_pt_ref + reset (nmpc_node.py:87-92)
            x1 = odom_to_x0(self.state[q])
            self.xr[q] = np.tile(x1, (CP.N_node + 1, 1))
            self.ur[q] = np.tile(np.array([0.0, 0.0, 0.0, CP.mass * CP.gravity]), (CP.N_node, 1))
            self.ctl[q].reset(self.xr[q], self.ur[q])

    def start_tracking(self):
        """pt_pub_callback (nmpc_node.py:135-152): the leader receives the trajectory, its reference becomes the first
        horizon and its controller is reset; its hover-throttle filter stops.  (The followers have no trajectory server;
        their filters keep running -- has_traj_server=False, nmpc_follower_node.py:28.)"""
        self.t_traj = 0.0
        self.xr[0], self.ur[0] = self.leader_reference(0.0)
        self.ctl[0].reset(self.xr[0], self.ur[0])

    def _control_tick(self):
        tracking = self.t_traj is not None
        if tracking:
            self.xr[0], self.ur[0] = self.leader_reference(self.t_traj)
            self.t_traj += CP.ts_nmpc
        u = [None] * 3
        tick = self.i // 2
        for q in range(3):
            if tick < self.START_TICK[q]:
                continue
            if q >= 1:
                # the followers consume the PredXU the leader published at the end of its tick (do_pub_ref, nmpc_node.py:229-230)
                self.xr[q], self.ur[q] = follower_reference(self.xr[0], self.ur[0], self.filters[q - 1].value)
            x0 = odom_to_x0(self.state[q])
            # hover-throttle estimation (50 Hz, same period as the control timer): vz from odometry, last thrust command
            if not (q == 0 and tracking):
                self.k_throttle[q] = self.est[q].update(float(self.state[q, 15]), float(self.thrust[q]))[0]
            if q == 0:
                f = self.disturb_force
                u[q] = self.ctl[q].update(x0, self.xr[q], self.ur[q], f)
            else:
                f = None
                u[q] = self.ctl[q].update(x0, self.xr[q], self.ur[q])
            if self.log is not None:
                self.log.append(dict(q=q, x0=x0, xr=self.xr[q].copy(), ur=self.ur[q].copy(), f=None if f is None else np.array(f, copy=True),
                                     u0=np.array(u[q], copy=True)))
            c = u0_to_cmd(np.asarray(u[q], dtype=np.float64), self.k_throttle[q])
            self.cmd[q] = c
            self.thrust[q] = float(c[3])
        # xiao_feng's PredXU reaches the leader (ndp_nmpc_leader_node.py:60-76): gate on the leader's odometry position
        if tick >= self.START_TICK[1]:
            self.disturb_force = leader_disturb_force(self.downwash_update, self.xr[1], self.state[0, 3:5], self.xr[0])
            if self.dw_log is not None and np.abs(self.disturb_force).max() > 0:
                self.dw_log.append(dict(other=self.xr[1].copy(), ego=self.xr[0].copy(), f=np.array(self.disturb_force, copy=True)))

    def sim_step(self):
        if self.i % 5 == 0:
            xf, sb = leader_formation_refs(float(self.state[0, 3]))
            self.filters[0].update(xf)
            self.filters[1].update(sb)
        if self.i % 2 == 0:
            self._control_tick()
        self.state = self.plant_forward(self.ts_sim, self.state, self.cmd)
        self.i += 1

    def run(self, control_ticks: int):
        for _ in range(2 * control_ticks):
            self.sim_step()
