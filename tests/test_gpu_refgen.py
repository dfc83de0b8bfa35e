"""GPU parity tests of the batched reference generation vs the reference's own functions
(golden fixture) and the numpy oracle."""
import numpy as np
import pytest

from conftest import golden
from ndp_nmpc_qd_b200 import traj_gen
from oracle import refgen_numpy as orf

pytestmark = pytest.mark.gpu
NAMES = ["eight_high_dyn", "eight_low", "eight_low_diff_h"]


def _golden_traj(g, name):
    m = len(g[name + "_t_cum"]) - 1
    return traj_gen.Trajectory(g[name + "_t_cum"], g[name + "_cx"].reshape(m, 8), g[name + "_cy"].reshape(m, 8), g[name + "_cz"].reshape(m, 8),
                               g[name + "_cyaw"].reshape(m, 4), g[name + "_wpts"][:, -1].copy())


def test_points_match_reference_functions(built_lib):
    """node 0 of the horizon at the golden sample times == get_traj_pt + diff_flatness of the reference
    (random times, exact knots, and the hover branch beyond the end); three trajectories in one table."""
    import torch
    from ndp_nmpc_qd_b200.traj_gen.refgen import RefGen

    g = golden("refgen_golden.npz")
    rg = RefGen([_golden_traj(g, n) for n in NAMES])
    for j, name in enumerate(NAMES):
        t = torch.as_tensor(g[name + "_t"], dtype=torch.float64, device="cuda")
        tid = torch.full((t.numel(),), j, dtype=torch.int32, device="cuda")
        xr, ur = rg.horizon(t, tid, N=1, dtype=torch.float64)
        assert np.abs(xr[:, 0].cpu().numpy() - g[name + "_x"]).max() < 1e-10, name
        assert np.abs(ur[:, 0].cpu().numpy() - g[name + "_u"]).max() < 1e-9, name


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 2e-6)])
def test_horizons_vs_oracle(built_lib, dtype, tol):
    import torch
    from ndp_nmpc_qd_b200.traj_gen.refgen import RefGen

    trs = [traj_gen.plan_named(n) for n in NAMES]
    rg = RefGen(trs)
    rng = np.random.default_rng(3)
    B = 257
    tid = rng.integers(0, 3, B).astype(np.int32)
    t0 = np.array([rng.uniform(-0.0, trs[j].duration + 1.0) for j in tid])
    off = rng.normal(size=(B, 3))
    dt = getattr(torch, dtype)
    xr, ur = rg.horizon(torch.as_tensor(t0, device="cuda"), torch.as_tensor(tid, device="cuda"), N=20, th_pred=0.1,
                        offset=torch.as_tensor(off, device="cuda"), dtype=dt)
    xr, ur = xr.cpu().numpy().astype(np.float64), ur.cpu().numpy().astype(np.float64)
    for b in range(B):
        xo, uo = orf.horizon(trs[tid[b]], t0[b], 20, 0.1, off[b])
        assert np.abs(xr[b] - xo).max() / max(1.0, np.abs(xo).max()) < tol, b
        assert np.abs(ur[b] - uo).max() / max(1.0, np.abs(uo).max()) < tol, b


def test_empty_and_long_horizon(built_lib):
    import torch
    from ndp_nmpc_qd_b200.traj_gen.refgen import RefGen

    rg = RefGen([traj_gen.plan_named("eight_low")])
    xr, ur = rg.horizon(torch.empty((0,), dtype=torch.float64, device="cuda"))
    assert xr.shape == (0, 21, 10) and ur.shape == (0, 20, 4)
    xr, ur = rg.horizon(torch.zeros((3,), dtype=torch.float64, device="cuda"), N=80)
    assert xr.shape == (3, 81, 10) and torch.isfinite(xr).all() and torch.isfinite(ur).all()
