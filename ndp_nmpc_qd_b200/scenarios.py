"""BASELINE.json's two reference-runnable configurations as deterministic, ROS-free scenario drivers.

  config 1 `one_qd_nmpc`        OneQuadTracking   (ndp_nmpc/launch/one_qd_nmpc.launch, dop_sim/config/one_qd_config.yaml)
  config 2 `three_qd_ndp_nmpc`  ThreeQuadFormation (ndp_nmpc_qd_b200/formation.py)

Both replace the ROS timers by a fixed schedule on the simulator's 0.01 s time base and take the controller, the
estimator and the plant as injected objects with the reference's interfaces, so the same scenario runs on the
drop-in CUDA classes and on the CPU oracles.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .formation import ThreeQuadFormation, odom_to_x0, u0_to_cmd  # noqa: F401
from .params import estimator_params as EP
from .params import nmpc_params as CP


class OneQuadTracking:
    """ControllerNode + DopQdNode for one quadrotor (nmpc_node.py:37-283, dop_qd_node.py:26-225).

    controller: `reset(xr, ur)`, `update(x0, xr, ur) -> u0`; plant_forward(ts_sim, state[1,35], cmd[1,4]) -> state;
    reference(t) -> (xr[21,10], ur[20,4]) at trajectory time t; estimator: HoverThrottleEstimator interface.
    Schedule per 0.01 s sim step i: [i % 2 == 0] control tick (hover-throttle update while not tracking, nmpc_node.py:
    146,196,251-253; x0 from odometry; controller.update; AttitudeTarget), then one plant step."""

    INIT_POS = np.array([[1.0, 1.0, 5.0]])  # one_qd_config.yaml:13-17

    def __init__(self, controller, plant_forward: Callable, reference: Callable, estimator, ts_sim: float = 0.01, record: bool = False):
        self.ctl, self.plant_forward, self.reference, self.est = controller, plant_forward, reference, estimator
        self.ts_sim = ts_sim
        self.k_throttle = EP.k_throttle_init
        self.state = np.zeros((1, 35))
        self.state[:, 9] = -1.0     # dop_qd_node.py:174
        self.state[:, 31:35] = 8.0  # dop_qd_node.py:175
        self.state[:, 3:6] = self.INIT_POS
        self.cmd = np.zeros((1, 4))
        self.cmd[:, 3] = 0.283      # dop_qd_node.py:190
        self.thrust = 0.0
        self.i = 0
        self.t_traj: Optional[float] = None
        self.log = [] if record else None
        x1 = odom_to_x0(self.state[0])  # gen_fix_pt_ref (pt_publisher.py:40-55) + reset (nmpc_node.py:87-92)
        self.xr = np.tile(x1, (CP.N_node + 1, 1))
        self.ur = np.tile(np.array([0.0, 0.0, 0.0, CP.mass * CP.gravity]), (CP.N_node, 1))
        self.ctl.reset(self.xr, self.ur)
        self.u0 = None

    def start_tracking(self):
        """pt_pub_callback (nmpc_node.py:135-152)."""
        self.t_traj = 0.0
        self.xr, self.ur = self.reference(0.0)
        self.ctl.reset(self.xr, self.ur)

    def _control_tick(self):
        tracking = self.t_traj is not None
        if tracking:
            self.xr, self.ur = self.reference(self.t_traj)
            self.t_traj += CP.ts_nmpc
        else:
            self.k_throttle = self.est.update(float(self.state[0, 15]), float(self.thrust))[0]
        x0 = odom_to_x0(self.state[0])
        self.u0 = self.ctl.update(x0, self.xr, self.ur)
        if self.log is not None:
            self.log.append(dict(x0=x0, xr=self.xr.copy(), ur=self.ur.copy(), u0=np.array(self.u0, copy=True), pos=self.state[0, 3:6].copy()))
        self.cmd[0] = u0_to_cmd(np.asarray(self.u0, dtype=np.float64), self.k_throttle)
        self.thrust = float(self.cmd[0, 3])

    def sim_step(self):
        if self.i % 2 == 0:
            self._control_tick()
        self.state = self.plant_forward(self.ts_sim, self.state, self.cmd)
        self.i += 1

    def run(self, control_ticks: int):
        for _ in range(2 * control_ticks):
            self.sim_step()
