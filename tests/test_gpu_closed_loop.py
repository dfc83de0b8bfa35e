"""GPU test of the device-resident closed loop (BASELINE.json config 5 / config 1 wiring): RefGen -> Engine ->
plant, against the same loop built from the three CPU oracles (fp64)."""
import numpy as np
import pytest

from ndp_nmpc_qd_b200 import traj_gen
from oracle import refgen_numpy as orf
from oracle.c_oracle import make_cfg
from oracle.plant_numpy import PlantOracle

pytestmark = pytest.mark.gpu


def _oracle_loop(c_oracle, trs, tid, t0, steps, ts_sim=0.01, ts_ctl_plant=0.01, k_thr=None):
    from ndp_nmpc_qd_b200.closed_loop import K_THROTTLE, MASS, GRAVITY

    k_thr = K_THROTTLE if k_thr is None else k_thr
    B = len(tid)
    hor = lambda t: [np.stack(a) for a in zip(*[orf.horizon(trs[tid[b]], t[b], 20, 0.1) for b in range(B)])]
    t = np.array(t0, dtype=np.float64)
    xr, ur = hor(t)
    s = np.zeros((B, 35))
    s[:, 3:6], s[:, 13:16], s[:, 9:13] = xr[:, 0, 0:3], xr[:, 0, 3:6], xr[:, 0, 6:10]
    s[:, 31:35] = np.sqrt(MASS * GRAVITY / 4 / (2.8158e-08 * 1e6))
    plant = PlantOracle(B, ts_sim, ts_ctl_plant, False, True, False)
    X, U = xr.copy(), ur.copy()
    cfg = make_cfg()
    u_hist = []
    for _ in range(steps):
        xr, ur = hor(t)
        x0 = np.concatenate([s[:, 3:6], s[:, 13:16], s[:, 9:13]], 1)
        r = c_oracle.rti_batch(cfg, x0, xr, ur, None, X, U)
        assert (r["status"] == 0).all()
        u0 = r["u0"]
        cmd = np.concatenate([u0[:, 0:3], u0[:, 3:4] * MASS / k_thr], 1)
        for _ in range(2):
            s = plant.forward(ts_sim, s, cmd)
        t = t + 0.02
        u_hist.append(u0.copy())
    return s, np.stack(u_hist)


@pytest.mark.parametrize("precision,tol_u,tol_p", [("f64", 1e-7, 1e-8), ("f32", 2e-3, 2e-4)])
def test_closed_loop_vs_oracles(built_lib, c_oracle, precision, tol_u, tol_p):
    import torch
    from ndp_nmpc_qd_b200.closed_loop import ClosedLoop

    trs = [traj_gen.plan_named("eight_high_dyn"), traj_gen.plan_named("eight_low_diff_h")]
    tid = np.array([0, 0, 0, 1, 1, 0], dtype=np.int32)
    t0 = np.array([0.0, 3.3, 11.7, 5.0, 20.0, 22.8])
    steps = 30
    cl = ClosedLoop(trs, tid, t0, precision=precision)
    u_hist = []
    for _ in range(steps):
        cl.step()
        u_hist.append(cl.u0.cpu().numpy().astype(np.float64))
    torch.cuda.synchronize()
    assert (cl.engine.status().cpu().numpy() == 0).all()
    s_ref, u_ref = _oracle_loop(c_oracle, trs, tid, t0, steps)
    s = cl.state.cpu().numpy()[:, :, 0]
    assert np.abs(np.stack(u_hist) - u_ref).max() < tol_u * max(1.0, np.abs(u_ref).max())
    assert np.abs(s[:, 3:6] - s_ref[:, 3:6]).max() < tol_p
    # tracking sanity: the quads stay on their references (plant = dop_sim model with rate-loop and motor lag)
    assert float(cl.position_error().max()) < 0.5


def test_closed_loop_large_batch_runs(built_lib):
    import torch
    from ndp_nmpc_qd_b200.closed_loop import ClosedLoop

    tr = traj_gen.plan_named("eight_high_dyn")
    B = 5000
    rng = np.random.default_rng(0)
    cl = ClosedLoop([tr], np.zeros(B, np.int32), rng.uniform(0, tr.duration - 3.0, B), offset=rng.normal(size=(B, 3)))
    for _ in range(50):
        cl.step()
    torch.cuda.synchronize()
    assert (cl.engine.status().cpu().numpy() == 0).all()
    err = cl.position_error().cpu().numpy()
    assert np.isfinite(err).all() and err.max() < 1.0 and np.sqrt((err**2).mean()) < 0.3
