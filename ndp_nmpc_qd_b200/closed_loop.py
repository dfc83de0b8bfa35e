"""Closed-loop batched rollouts entirely on the device (BASELINE.json config 5): for B independent
scenarios, per 50 Hz control step
    reference horizon (RefGen, pt_pub)  ->  controller.update (Engine: one SQP-RTI step, [+ downwash MLP])
    ->  AttitudeTarget mapping (nmpc_node.py:273-283)  ->  plant steps (MulQuadrotors, dop_sim)  ->  odometry -> x0
mirroring the ROS wiring nmpc_node.py <-> dop_qd_node.py without the host in
the loop.  Everything is a kernel of libndp_nmpc_b200.so; torch only owns the buffers.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .dop_sim import MulQuadrotors
from .params import nmpc_params as CP
from .solver import Engine
from .traj_gen.min_snap import Trajectory
from .traj_gen.refgen import RefGen
from .wire import BatchedHoverThrottleEstimator

MASS, GRAVITY = 1.4844, 9.81
# throttle that makes the plant hover: 4 (k_th thr + b_th) = m g (atp_rate.py:92, control_param.py:19-20);
# k_throttle = m g / thr_hover is what the reference's HoverThrottleEstimator converges to
THR_HOVER = (MASS * GRAVITY / 4 + 1.206) / 17.666
K_THROTTLE = MASS * GRAVITY / THR_HOVER


class ClosedLoop:
    def __init__(self, trajectories: Sequence[Trajectory], traj_id: np.ndarray, t_start: np.ndarray, N: int = CP.N_node,
                 precision: str = "f32", ts_sim: float = 0.01, ts_ctl_plant: float = 0.01, ts_nmpc: float = CP.ts_nmpc,
                 offset: Optional[np.ndarray] = None, has_motor_model: bool = True, has_battery: bool = False,
                 has_downwash: bool = False, group: int = 1, k_throttle: Optional[float] = K_THROTTLE, device="cuda:0",
                 downwash_mlp: bool = False, **engine_overrides):
        """k_throttle: fixed thrust scale of nmpc_u_2_att_tgt, or None to run the reference's HoverThrottleEstimator on the
        device (one filter per scenario, estimator_params.py:13 initial guess 50): like the node, the filter is updated
        only while no trajectory is tracked (`hover()`; nmpc_node.py:146,196) and its estimate is frozen during `step()`.
        downwash_mlp: the coupled variant -- scenarios are contiguous blocks of `group` quadrotors (formation offsets via
        `offset`), the plant couples them physically (has_downwash, same blocks) and every controller is the NDP-NMPC one:
        per step the forces are the gated sum of DownwashNN over the block's other reference horizons
        (ndp_nmpc_leader_node.py:60-76, generalised to every quad of the block; gate on the ego's odometry position)."""
        self.device = torch.device(device)
        B = len(traj_id)
        self.B, self.N, self.ts_sim, self.ts_nmpc = B, N, ts_sim, ts_nmpc
        self.sim_per_ctl = int(round(ts_nmpc / ts_sim))
        self.k_throttle = None if k_throttle is None else float(k_throttle)
        self.dtype = torch.float32 if precision == "f32" else torch.float64
        dev = self.device
        self.refgen = RefGen(trajectories, device=dev)
        self.downwash_mlp = bool(downwash_mlp)
        self.engine = Engine(batch=B, N=N, np_=7 if downwash_mlp else 4, precision=precision, device=dev, **engine_overrides)
        self.f = None
        if downwash_mlp:
            from .dnwash_nn_est import DownwashNN

            self.nn = DownwashNN(device=dev)
            self.nn.set_group(group if 0 < group < B else 0)
            self.traj6 = torch.empty((B, N + 1, 6), dtype=torch.float32, device=dev)
            self.odom_xy = torch.empty((B, 2), dtype=torch.float32, device=dev)
        self.plant = MulQuadrotors(B, ts_sim, ts_ctl_plant, torch.float64, has_downwash, has_motor_model, has_battery, group=group, device=dev)
        self.traj_id = torch.as_tensor(np.asarray(traj_id, dtype=np.int32), device=dev)
        self.t = torch.as_tensor(np.asarray(t_start, dtype=np.float64), device=dev)
        self.offset = None if offset is None else torch.as_tensor(np.asarray(offset, dtype=np.float64), device=dev).contiguous()
        self.xr = torch.empty((B, N + 1, 10), dtype=self.dtype, device=dev)
        self.ur = torch.empty((B, N, 4), dtype=self.dtype, device=dev)
        self.x0 = torch.empty((B, 10), dtype=self.dtype, device=dev)
        self.u0 = torch.empty((B, 4), dtype=self.dtype, device=dev)
        self.cmd = torch.zeros((B, 4, 1), dtype=torch.float64, device=dev)
        self.state = torch.zeros((B, 35, 1), dtype=torch.float64, device=dev)
        self.estimator = BatchedHoverThrottleEstimator(B, ts_nmpc, device=dev) if k_throttle is None else None
        self.reset_to_reference()

    def reset_to_reference(self):
        """plant on the reference at t (rotors at hover speed), iterate = reference horizon (controller.reset)."""
        self.refgen.horizon(self.t, self.traj_id, self.N, CP.th_pred, self.offset, xr=self.xr, ur=self.ur)
        s = self.state
        s.zero_()
        x = self.xr[:, 0].to(torch.float64)
        s[:, 3:6, 0] = x[:, 0:3]; s[:, 13:16, 0] = x[:, 3:6]; s[:, 9:13, 0] = x[:, 6:10]
        s[:, 31:35, 0] = float(np.sqrt(MASS * GRAVITY / 4 / (2.8158e-08 * 1e6)))  # hover rotor speed [kRPM]
        self.plant.reset()
        self.engine.reset(self.xr, self.ur)

    def step(self):
        """one control period: returns nothing; self.u0 / self.state hold the latest command and plant state."""
        self.refgen.horizon(self.t, self.traj_id, self.N, CP.th_pred, self.offset, xr=self.xr, ur=self.ur)
        self.plant.nmpc_x0(self.state, self.x0)
        if self.downwash_mlp:
            self.traj6.copy_(self.xr[:, :, 0:6])
            self.odom_xy.copy_(self.state[:, 3:5, 0])
            self.f = self.nn.forward_swarm(self.traj6, 0, self.B, odom_xy=self.odom_xy, out_dtype=self.dtype)
        self.engine.update(self.x0, self.xr, self.ur, self.f, self.u0)
        self._command_and_simulate()
        self.t.add_(self.ts_nmpc)

    def _command_and_simulate(self):
        if self.estimator is not None:
            self.plant.cmd_from_u0_dev(self.u0, self.cmd, MASS, self.estimator.k_throttle)
        else:
            self.plant.cmd_from_u0(self.u0, self.cmd, MASS, self.k_throttle)
        for _ in range(self.sim_per_ctl):
            self.plant(self.ts_sim, self.state, self.cmd)

    def hover(self, steps: int):
        """The node before a trajectory arrives (nmpc_node.py:84-101): the reference is the fixed point of
        gen_fix_pt_ref (pt_publisher.py:40-55: every node = the odometry at start-up, u_ref = (0, 0, 0, m g) -- the
        reference's own quirk), the controller runs at 50 Hz and the hover-throttle filter is
        updated every tick from (odometry v_z, last thrust command)."""
        x_fix = torch.empty((self.B, 10), dtype=self.dtype, device=self.device)
        self.plant.nmpc_x0(self.state, x_fix)
        self.xr.copy_(x_fix[:, None, :].expand(-1, self.N + 1, -1))
        self.ur.zero_()
        self.ur[:, :, 3] = MASS * GRAVITY
        self.engine.reset(self.xr, self.ur)
        for _ in range(steps):
            if self.estimator is not None:
                self.estimator.update(self.state[:, 15, 0], self.cmd[:, 3, 0])
            self.plant.nmpc_x0(self.state, self.x0)
            self.engine.update(self.x0, self.xr, self.ur, None, self.u0)
            self._command_and_simulate()

    def position_error(self) -> torch.Tensor:
        """|p - p_ref(t)| per scenario, against the reference at the current time."""
        xr, _ = self.refgen.horizon(self.t, self.traj_id, 1, CP.th_pred, self.offset, dtype=torch.float64)
        return (self.state[:, 3:6, 0] - xr[:, 0, 0:3]).norm(dim=1)

    @property
    def launch_count(self) -> int:
        return self.engine.launch_count + self.plant.launch_count + self.refgen.launch_count + (self.nn.launch_count if self.downwash_mlp else 0)


def time_closed_loop(N: int, B: int, steps: int = 250, precision: str = "f32", device="cuda:0", seed: Optional[int] = None, use_graph: bool = True,
                     trajectories=None) -> dict:
    """One point of BASELINE.json config 5 (horizon x batch sweep, closed loop with batched dop_sim rollouts) on one
    GPU: B scenarios on randomly phased eight_high_dyn / eight_low references, `steps` control steps (two plant steps
    each) timed with CUDA events.  The control step (6 library kernels + a clock update) is captured once into a CUDA
    graph and replayed, so small batches measure the device and not the Python launch rate."""
    from . import traj_gen

    dev = torch.device(device)
    trs = trajectories or [traj_gen.plan_named("eight_high_dyn"), traj_gen.plan_named("eight_low")]
    rng = np.random.default_rng(B + N if seed is None else seed)
    tid = (rng.random(B) < 0.5).astype(np.int32)
    t0 = np.array([rng.uniform(0, trs[j].duration - 6.0) for j in tid])
    with torch.cuda.device(dev):
        cl = ClosedLoop(trs, tid, t0, N=N, precision=precision, offset=rng.normal(size=(B, 3)) * 2.0, device=dev)
        for _ in range(5):
            cl.step()
        torch.cuda.synchronize(dev)
        graph, mode = None, "eager"
        if use_graph:
            try:
                cap = torch.cuda.Stream(device=dev)
                cap.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(cap):
                    cl.step()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=cap):
                        cl.step()
                torch.cuda.current_stream(dev).wait_stream(cap)
                graph, mode = g, "cuda graph replay"
            except Exception:  # noqa: BLE001
                graph = None
        torch.cuda.synchronize(dev)
        l0 = cl.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            if graph is not None:
                graph.replay()
            else:
                cl.step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        launches = (cl.launch_count - l0) / steps if graph is None else None
        sq = torch.zeros((), dtype=torch.float64, device=dev)
        n_err = min(steps, 50)
        for _ in range(n_err):
            cl.step()
            sq += (cl.position_error() ** 2).mean()
        torch.cuda.synchronize(dev)
        st = cl.engine.status().cpu().numpy()
        stats = cl.engine.stats().cpu().numpy()
        out = dict(N=N, batch=B, control_steps=steps, ms_per_control_step=ms / steps, solves_per_s=B * steps / (ms * 1e-3),
                   sim_steps_per_s=cl.sim_per_ctl * B * steps / (ms * 1e-3), pos_rmse_m=float(torch.sqrt(sq / n_err)),
                   status_nonzero=int((st != 0).sum()), riccati_sweeps_mean=float(stats[:, 0].mean()), launch_mode=mode, precision=precision)
        if launches is not None:
            out["launches_per_step"] = launches
        del cl
        torch.cuda.empty_cache()
    return out
