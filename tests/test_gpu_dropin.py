"""GPU tests of the drop-in controller surface (batch 1), mirroring how nmpc_node.py drives the
reference classes: ctor -> reset(xr, ur) -> update(x0, xr, ur[, f]) -> u0; solver.get(i, "x")."""
import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import mlp_numpy
from oracle.c_oracle import make_cfg
from ndp_nmpc_qd_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def test_nmpc_controller_closed_loop_vs_oracle(built_lib, c_oracle):
    """config 1 (one_qd_nmpc): track the high-dynamics eight for 40 control steps of 0.02 s with the
    nominal model as plant; u0 of every step within 1e-4 of the oracle driven identically."""
    from ndp_nmpc_qd_b200.nmpc_ctl import NMPCBodyRateController
    from oracle import nmpc_numpy as on

    p = on.OcpParams(T=0.02, N=1)  # plant step = one RK4 step of 0.02 s
    ctl = NMPCBodyRateController(is_build_acados=False)
    assert ctl.solver.N == 20
    t = 1.0
    xr, ur = wl.reference_horizon([t])
    x = xr[0, 0].copy()
    ctl.reset(xr[0], ur[0])
    oX, oU = xr.copy(), ur.copy()
    cfg = make_cfg()
    for step in range(40):
        xr, ur = wl.reference_horizon([t])
        u0 = ctl.update(x, xr[0], ur[0])
        r = c_oracle.rti_batch(cfg, x[None], xr, ur, None, oX, oU)
        assert u0.shape == (4,) and u0.dtype == np.float64
        assert rel_err(u0[None], r["u0"]) < 1e-4, step
        xs = ctl.solver.get(5, "x")
        assert xs.shape == (10,) and rel_err(xs[None], oX[:, 5]) < 1e-4
        xs[:] = 0  # the node mutates what get() returns (nmpc_node.py:237-238): must be a copy
        assert np.abs(ctl.solver.get(5, "x")).max() > 0
        x, _, _ = on.rk4_sens(x, r["u0"][0], np.zeros(3), p)
        x[6:10] /= np.linalg.norm(x[6:10])
        t += 0.02
    assert ctl.solver.status == 0


def test_ndp_controller_with_downwash(built_lib, c_oracle, mlp_weights):
    """config 2 leader: DownwashNN.update -> NDPNMPCBodyRateController.update(x0, xr, ur, f)."""
    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.ndp_nmpc_ctl import NDPNMPCBodyRateController
    from ndp_nmpc_qd_b200.nmpc_ctl import NMPCBodyRateController

    g = golden("downwash_golden.npz")
    nn = DownwashNN()
    f = nn.update(g["other"], g["ego"])
    assert f.shape == (21, 3) and f.dtype == np.float32
    assert np.abs(f - g["f"]).max() < 1e-5  # vs the reference's torch module (golden)
    ctl = NDPNMPCBodyRateController()
    assert not isinstance(ctl, NMPCBodyRateController)  # siblings, like the reference's classes (nmpc_node.py:203-208)
    xr, ur = wl.reference_horizon([3.0], name="eight_low")
    ctl.reset(xr[0], ur[0])
    x0 = xr[0, 0] + np.array([0.05, -0.03, 0.02, 0.1, 0, -0.1, 0, 0, 0, 0])
    u0 = ctl.update(x0, xr[0], ur[0], f)
    X, U = xr.copy(), ur.copy()
    r = c_oracle.rti_batch(make_cfg(), x0[None], xr, ur, f[None].astype(np.float64), X, U)
    assert rel_err(u0[None], r["u0"]) < 1e-4
    # the force matters: without it the collective acceleration differs
    ctl2 = NDPNMPCBodyRateController()
    ctl2.reset(xr[0], ur[0])
    u0z = ctl2.update(x0, xr[0], ur[0], np.zeros((21, 3)))
    assert abs(u0z[3] - u0[3]) > 1e-2


def test_acados_style_set_get_surface(built_lib, c_oracle):
    """The 42 per-stage solver.set calls exactly as the reference writes them
    (nmpc_body_rate_ctl.py:93-107)."""
    from ndp_nmpc_qd_b200.solver import BatchedOcpSolver

    s = BatchedOcpSolver(np_=7)
    xr, ur = wl.reference_horizon([5.0])
    xr, ur = xr[0], ur[0]
    f = np.random.default_rng(0).normal(size=(21, 3))
    for i in range(s.N):
        s.set(i, "x", xr[i, :])
        s.set(i, "u", ur[i, :])
    s.set(s.N, "x", xr[s.N, :])
    for i in range(s.N):
        s.set(i, "yref", np.concatenate((xr[i, :], ur[i, :])))
        s.set(i, "p", np.concatenate((xr[i, 6:10], f[i, :])))
    s.set(s.N, "yref", xr[s.N, :])
    s.set(s.N, "p", np.concatenate((xr[s.N, 6:10], f[s.N, :])))
    x0 = xr[0] + 0.01
    u0 = s.solve_for_x0(x0)
    assert s.status == 0
    X, U = xr[None].copy(), ur[None].copy()
    r = c_oracle.rti_batch(make_cfg(), x0[None], xr[None], ur[None], f[None], X, U)
    assert rel_err(u0[None], r["u0"]) < 1e-4
    assert rel_err(s.get(0, "u")[None], U[:, 0]) < 1e-4
    assert rel_err(s.get(s.N, "x")[None], X[:, s.N]) < 1e-4
    with pytest.raises(Exception):
        s.set(0, "nope", x0)


@pytest.mark.parametrize("B", [7, 1000])
def test_batched_mirror_equals_device_path(built_lib, B):
    """BatchedOcpSolver with a batch axis (host mirrors, one ndp_solve_host call per solve: zero-copy x0 / u0 for the
    small batch, staged copies for the large one) == the device-pointer Engine path, three warm-started steps."""
    import torch

    from ndp_nmpc_qd_b200.solver import BatchedOcpSolver, Engine

    w = wl.independent_problems(B, seed=90 + B, scale=3.0)
    fd = np.random.default_rng(3).normal(size=(B, 21, 3))
    s = BatchedOcpSolver(batch=B, np_=7)
    s.reset(w["xr"], w["ur"])
    e = Engine(batch=B, np_=7)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    xr, ur, f = t(w["xr"]), t(w["ur"]), t(fd)
    e.reset(xr, ur)
    e.set_reference(xr, ur, f)
    rng = np.random.default_rng(4)
    for step in range(3):
        x0 = w["x0"] + 0.02 * step * rng.normal(size=w["x0"].shape)
        if step != 1:   # step 1 reuses the uploaded reference (ndp_solve_host without the upload)
            s.set_reference(w["xr"], w["ur"], fd)
        u0 = s.solve_for_x0(x0)
        ref = e.solve(t(x0)).cpu().numpy()
        assert np.array_equal(u0.astype(np.float32), ref), step
        assert np.array_equal(s.status, e.status().cpu().numpy())
    assert np.array_equal(s.get_all("x").astype(np.float32), e.get_all("x").cpu().numpy())
