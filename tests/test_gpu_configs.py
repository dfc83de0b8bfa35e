"""BASELINE.json configs 1 and 2 in full, drop-in CUDA classes vs the CPU oracles.

  config 1 `one_qd_nmpc`        250 hover ticks + all 1 157 control ticks of eight_high_dyn.yaml
  config 2 `three_qd_ndp_nmpc`  150 hover ticks + all 1 789 control ticks of eight_low.yaml, leader with DownwashNN +
                                two followers with the alpha-filtered formation offsets and the offset switch

Each scenario (ndp_nmpc_qd_b200/scenarios.py, formation.py) is run once on the oracle stack (C SQP-RTI oracle, numpy
MLP, numpy dop_sim plant, all fp64) with every controller call logged.  Then
  (replay)   the drop-in controllers are fed the logged (x0, xr, ur, f) tick by tick -- their own warm-started iterate
             against the oracle's over the whole trajectory -- and every u0 must agree per component;
  (free run) the whole scenario is run again on the CUDA classes (controllers, DownwashNN, MulQuadrotors plant) and the
             closed-loop trajectories must coincide.
Gate: |u0_i - ref_i| <= 1e-4 max(|ref_i|, 1) per component (north_star: 1e-4 relative in fp32).
"""
import numpy as np
import pytest

from helpers.oracle_backends import OracleController, OraclePlant, leader_reference, oracle_downwash_update
from ndp_nmpc_qd_b200 import scenarios as sc
from ndp_nmpc_qd_b200 import traj_gen
from ndp_nmpc_qd_b200.hv_throttle_est import HoverThrottleEstimator

pytestmark = pytest.mark.gpu

TOL = 1e-4


def comp_err(u, ref):
    u, ref = np.asarray(u, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(u - ref) / np.maximum(np.abs(ref), 1.0)))


class CudaPlant:
    """MulQuadrotors (CUDA) behind the numpy interface of the scenario drivers."""

    def __init__(self, n):
        import torch
        from ndp_nmpc_qd_b200.dop_sim import MulQuadrotors

        self.torch = torch
        self.m = MulQuadrotors(n, 0.01, 0.01, torch.float64, True, True, True)
        self.s = torch.zeros((n, 35, 1), dtype=torch.float64, device="cuda")
        self.c = torch.zeros((n, 4, 1), dtype=torch.float64, device="cuda")

    def forward(self, ts_sim, state, cmd):
        t = self.torch
        self.s.copy_(t.from_numpy(np.ascontiguousarray(state))[:, :, None])
        self.c.copy_(t.from_numpy(np.ascontiguousarray(cmd))[:, :, None])
        self.m.forward(ts_sim, self.s, self.c)
        return self.s[:, :, 0].cpu().numpy()


@pytest.fixture(scope="module")
def config1_oracle(c_oracle):
    tr = traj_gen.plan_named("eight_high_dyn")
    s = sc.OneQuadTracking(OracleController(c_oracle), OraclePlant(1).forward, leader_reference(tr), HoverThrottleEstimator(0.02), record=True)
    s.run(250)
    s.start_tracking()
    n_track = int(np.ceil(tr.duration / 0.02))
    assert n_track == 1157
    s.run(n_track)
    return tr, s


def test_config1_full_trajectory_replay(built_lib, config1_oracle):
    from ndp_nmpc_qd_b200.nmpc_ctl import NMPCBodyRateController

    tr, s = config1_oracle
    ctl = NMPCBodyRateController(is_build_acados=True)
    worst, sat = 0.0, 0
    for k, rec in enumerate(s.log):
        if k == 0 or k == 250:  # gen_fix_pt_ref reset / pt_pub_callback reset
            ctl.reset(rec["xr"], rec["ur"])
        u0 = ctl.update(rec["x0"], rec["xr"], rec["ur"])
        e = comp_err(u0, rec["u0"])
        assert e < TOL, (k, e, u0, rec["u0"])
        worst = max(worst, e)
        sat += int(np.any(np.abs(rec["u0"][:3]) >= 6.0 - 1e-9))
    print(f"config 1 replay: {len(s.log)} ticks, worst per-component u0 error {worst:.2e}, ticks with a saturated body rate {sat}")
    assert len(s.log) == 250 + 1157 and sat > 0  # the dop_sim plant drives this config into its input bounds


def test_config1_free_run(built_lib, config1_oracle):
    from ndp_nmpc_qd_b200.nmpc_ctl import NMPCBodyRateController

    tr, s = config1_oracle
    g = sc.OneQuadTracking(NMPCBodyRateController(), CudaPlant(1).forward, leader_reference(tr), HoverThrottleEstimator(0.02), record=True)
    g.run(250)
    g.start_tracking()
    g.run(1157)
    pos_g, pos_o = np.array([r["pos"] for r in g.log]), np.array([r["pos"] for r in s.log])
    err = lambda log: np.sqrt(np.mean([np.sum((r["pos"] - r["xr"][0, :3]) ** 2) for r in log[250:]]))
    print(f"config 1 free run: max |p_gpu - p_oracle| {np.abs(pos_g - pos_o).max():.2e} m, tracking RMSE gpu {err(g.log):.5f} oracle {err(s.log):.5f}")
    assert np.abs(pos_g - pos_o).max() < 2e-3
    assert abs(err(g.log) - err(s.log)) < 1e-4
    assert abs(g.k_throttle - s.k_throttle) < 1e-4


@pytest.fixture(scope="module")
def config2_oracle(c_oracle, mlp_weights):
    tr = traj_gen.plan_named("eight_low")
    f = sc.ThreeQuadFormation(OracleController(c_oracle, True), [OracleController(c_oracle), OracleController(c_oracle)],
                              oracle_downwash_update(mlp_weights), OraclePlant(3).forward, leader_reference(tr),
                              lambda: HoverThrottleEstimator(0.02), record=True)
    f.run(150)
    f.start_tracking()
    n_track = int(np.ceil(tr.duration / 0.02))
    assert n_track == 1789
    f.run(n_track)
    return tr, f


def test_config2_formation_replay(built_lib, config2_oracle):
    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.ndp_nmpc_ctl import NDPNMPCBodyRateController
    from ndp_nmpc_qd_b200.nmpc_ctl import NMPCBodyRateController

    tr, f = config2_oracle
    ctl = [NDPNMPCBodyRateController(), NMPCBodyRateController(is_build_acados=False), NMPCBodyRateController(is_build_acados=False)]
    nn = DownwashNN()
    seen = [0, 0, 0]
    worst, worst_f, n_force = 0.0, 0.0, 0
    track_first = None
    for rec in f.log:
        q = rec["q"]
        if seen[q] == 0:
            ctl[q].reset(rec["xr"], rec["ur"])
        if q == 0 and track_first is None and not np.allclose(rec["xr"][0], rec["xr"][-1]):
            track_first = seen[0]
            ctl[0].reset(rec["xr"], rec["ur"])  # pt_pub_callback reset
        seen[q] += 1
        if q == 0:
            u0 = ctl[0].update(rec["x0"], rec["xr"], rec["ur"], rec["f"])
            n_force += int(np.abs(rec["f"]).max() > 0)
        else:
            u0 = ctl[q].update(rec["x0"], rec["xr"], rec["ur"])
        e = comp_err(u0, rec["u0"])
        assert e < TOL, (q, seen[q], e, u0, rec["u0"])
        worst = max(worst, e)
    # the downwash observer on the (other, ego) pairs that passed the leader's gate
    checked = 0
    for rec in f.dw_log[:: max(1, len(f.dw_log) // 200)]:
        worst_f = max(worst_f, float(np.abs(nn.update(rec["other"], rec["ego"]) - rec["f"]).max()))
        checked += 1
    print(f"config 2 replay: {len(f.log)} controller calls, worst per-component u0 error {worst:.2e}; leader ticks with a downwash force "
          f"{n_force}, DownwashNN.update vs oracle on {checked} of them: {worst_f:.2e} N")
    assert seen == [150 + 1789, 150 + 1789 - 3, 150 + 1789 - 7] and track_first == 150
    assert n_force > 500 and checked >= 200 and worst_f < 2e-5


def test_config2_formation_free_run(built_lib, config2_oracle):
    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.ndp_nmpc_ctl import NDPNMPCBodyRateController
    from ndp_nmpc_qd_b200.nmpc_ctl import NMPCBodyRateController

    tr, f = config2_oracle
    g = sc.ThreeQuadFormation(NDPNMPCBodyRateController(), [NMPCBodyRateController(is_build_acados=False), NMPCBodyRateController(is_build_acados=False)],
                              DownwashNN().update, CudaPlant(3).forward, leader_reference(tr), lambda: HoverThrottleEstimator(0.02), record=True)
    g.run(150)
    g.start_tracking()
    g.run(1789)
    assert len(g.log) == len(f.log)
    du = max(comp_err(a["u0"], b["u0"]) for a, b in zip(g.log, f.log))
    dp = float(np.abs(g.state[:, 3:6] - f.state[:, 3:6]).max())
    print(f"config 2 free run: worst u0 difference along the closed loops {du:.2e}, final position difference {dp:.2e} m, "
          f"filtered offsets {g.filters[0].value} {g.filters[1].value}")
    assert dp < 2e-3 and du < 5e-3
    assert np.allclose(g.filters[0].value, f.filters[0].value, atol=1e-9)
