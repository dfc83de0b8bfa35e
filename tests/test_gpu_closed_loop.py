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


def test_coupled_closed_loop_with_downwash_mlp(built_lib, c_oracle, mlp_weights):
    """Coupled variant of config 5: blocks of 3 quadrotors (a follower 0.6 m above its leader, one beside it) whose plants
    interact through the simulator's pairwise downwash and whose NDP-NMPC controllers get the gated DownwashNN sum of the
    block's other reference horizons.  Against the same loop on the CPU oracles (numpy MLP / swarm sum, C SQP-RTI, numpy
    plant); two blocks far apart in index but at the SAME place in space check that `group` keeps scenarios independent."""
    import torch
    from ndp_nmpc_qd_b200.closed_loop import ClosedLoop, K_THROTTLE, MASS, GRAVITY
    from oracle import mlp_numpy

    tr = traj_gen.plan_named("eight_low")
    G, steps = 3, 25
    off1 = np.array([[0.0, 0.0, 0.0], [0.05, 0.1, 0.6], [0.0, 1.2, 0.0]])
    off = np.concatenate([off1, off1])          # second block: same positions -> would couple without `group`
    tid = np.zeros(2 * G, np.int32)
    t0 = np.array([5.0] * G + [5.0] * G)
    cl = ClosedLoop([tr], tid, t0, precision="f64", offset=off, has_downwash=True, has_battery=False, group=G, downwash_mlp=True)
    u_hist, f_hist = [], []
    for _ in range(steps):
        cl.step()
        u_hist.append(cl.u0.cpu().numpy().copy()); f_hist.append(cl.f.cpu().numpy().copy())
    torch.cuda.synchronize()
    assert (cl.engine.status().cpu().numpy() == 0).all()
    u_hist, f_hist = np.stack(u_hist), np.stack(f_hist)
    assert np.array_equal(u_hist[:, :G], u_hist[:, G:])  # the two blocks are identical scenarios and do not see each other
    assert np.abs(f_hist[:, 0]).max() > 0.1 and np.abs(f_hist[:, 2]).max() == 0.0  # leader under the follower; the third outside the gate
    # oracle loop for one block
    hor = lambda t: [np.stack(a) for a in zip(*[orf.horizon(tr, t[b], 20, 0.1, off1[b]) for b in range(G)])]
    t = np.array(t0[:G], dtype=np.float64)
    xr, ur = hor(t)
    s = np.zeros((G, 35))
    s[:, 3:6], s[:, 13:16], s[:, 9:13] = xr[:, 0, 0:3], xr[:, 0, 3:6], xr[:, 0, 6:10]
    s[:, 31:35] = np.sqrt(MASS * GRAVITY / 4 / (2.8158e-08 * 1e6))
    plant = PlantOracle(G, 0.01, 0.01, True, True, False)
    X, U = xr.copy(), ur.copy()
    cfg = make_cfg()
    for k in range(steps):
        xr, ur = hor(t)
        x0 = np.concatenate([s[:, 3:6], s[:, 13:16], s[:, 9:13]], 1)
        f = mlp_numpy.swarm_forces(mlp_weights, xr[:, :, 0:6].astype(np.float32), 0, G, odom_xy=s[:, 3:5].astype(np.float32))
        r = c_oracle.rti_batch(cfg, x0, xr, ur, f, X, U)
        assert (r["status"] == 0).all()
        assert np.abs(f_hist[k, :G] - f).max() < 2e-4, k
        assert np.abs(u_hist[k, :G] - r["u0"]).max() < 2e-4 * max(1.0, np.abs(r["u0"]).max()), k  # forces from the fp32-accurate MLP
        cmd = np.concatenate([r["u0"][:, 0:3], r["u0"][:, 3:4] * MASS / K_THROTTLE], 1)
        for _ in range(2):
            s = plant.forward(0.01, s, cmd)
        t = t + 0.02
    assert np.abs(cl.state.cpu().numpy()[:G, 3:6, 0] - s[:, 3:6]).max() < 1e-4
