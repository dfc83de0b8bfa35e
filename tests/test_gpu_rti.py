"""GPU parity tests of the NMPC hot path: CUDA engine (through the C ABI) vs the fp64 CPU oracle.

Tolerances (BASELINE.json north_star / SURVEY.md 8d): u0 and predicted trajectories within 1e-4
relative for the fp32 build, 1e-9 for the fp64 build, measured per component:
max over all elements of |a - b| / max(|b|, 1) (conftest.rel_err)."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err
from oracle.c_oracle import make_cfg
from ndp_nmpc_qd_b200 import workloads as wl

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-4, "f64": 1e-9}


def _dt(prec):
    return torch.float32 if prec == "f32" else torch.float64


def _engine(B, prec, N=20, np_=7, **kw):
    from ndp_nmpc_qd_b200.solver import Engine

    return Engine(batch=B, N=N, np_=np_, precision=prec, **kw)


def _run_engine(e, w, fd=None, steps=1, x0_seq=None):
    dt, dev = e.dtype, e.device
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    xr, ur = t(w["xr"]), t(w["ur"])
    e.reset(xr, ur)
    e.set_reference(xr, ur, None if fd is None else t(fd))
    u0 = None
    for s in range(steps):
        x0 = t(w["x0"] if x0_seq is None else x0_seq[s])
        u0 = e.solve(x0)
    torch.cuda.synchronize()
    return (u0.cpu().numpy().astype(np.float64), e.get_all("x").cpu().numpy().astype(np.float64),
            e.get_all("u").cpu().numpy().astype(np.float64), e.status().cpu().numpy(), e.stats().cpu().numpy())


def _run_oracle(c_oracle, cfg, w, fd=None, steps=1, x0_seq=None):
    X, U = w["xr"].copy(), w["ur"].copy()
    r = None
    for s in range(steps):
        r = c_oracle.rti_batch(cfg, w["x0"] if x0_seq is None else x0_seq[s], w["xr"], w["ur"], fd, X, U)
    return r["u0"], X, U, r["status"], r


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_rk4_sens_kernel_vs_oracle(built_lib, c_oracle, prec):
    """(1) batched RK4 integrator with forward sensitivities."""
    import ctypes as C

    from ndp_nmpc_qd_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(0)
    M = 1000
    x = rng.normal(size=(M, 10)); x[:, 6:10] /= np.linalg.norm(x[:, 6:10], axis=1, keepdims=True)
    u = np.concatenate([rng.normal(size=(M, 3)) * 2, 9.81 + 3 * rng.normal(size=(M, 1))], 1)
    f = rng.normal(size=(M, 3)) * 2
    x[0], u[0], f[0] = [0.1, -0.2, 0.3, 1.0, -2.0, 0.5, 0.9, 0.1, -0.2, 0.3], [1.2, -0.7, 0.4, 13.0], [0.3, -0.2, -4.0]
    dt = _dt(prec)
    tx, tu, tf = (torch.as_tensor(a, dtype=dt, device="cuda") for a in (x, u, f))
    xn = torch.empty((M, 10), dtype=dt, device="cuda")
    AB = torch.empty((M, 10, 14), dtype=dt, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.ndp_rk4_sens(0 if prec == "f32" else 1, M, 0.1, 1.4844, 9.81, p(tx), p(tu), p(tf), p(xn), p(AB), None)
    assert rc == 0
    torch.cuda.synchronize()
    cfg = make_cfg()
    ref_x = np.zeros((M, 10)); ref_AB = np.zeros((M, 10, 14))
    xi, ui, fi = (t.cpu().numpy().astype(np.float64) for t in (tx, tu, tf))
    for m in range(M):
        a, Sx, Su = c_oracle.rk4_sens(cfg, xi[m], ui[m], fi[m])
        ref_x[m], ref_AB[m, :, :10], ref_AB[m, :, 10:] = a, Sx, Su
    tol = 1e-6 if prec == "f32" else 1e-12
    assert rel_err(xn.cpu().numpy(), ref_x) < tol
    assert rel_err(AB.cpu().numpy(), ref_AB) < tol
    if prec == "f64":  # SURVEY.md B.5 known answer
        assert abs(xn[0, 3].item() - 0.646352242513726) < 1e-13


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_golden_problems(built_lib, prec):
    """The committed dense-KKT solutions (tests/golden/rti_golden.npz), incl. up to 10 active bounds."""
    g = golden("rti_golden.npz")
    B = g["x0"].shape[0]
    e = _engine(B, prec)
    w = dict(x0=g["x0"], xr=g["xr"], ur=g["ur"])
    u0, X, U, st, stats = _run_engine(e, w, g["fd"])
    assert np.all(st == 0), st
    assert rel_err(u0, g["u0"]) < TOL[prec]
    assert rel_err(X, g["X"]) < TOL[prec] and rel_err(U, g["U"]) < TOL[prec]
    assert stats[:4, 0].max() == 1  # nominal problems: one Riccati sweep, no IPM
    assert stats[8:, 0].min() >= 2  # saturated problems needed constrained sweeps (active-set rounds or IPM)
    # the IPM-first route (active_set_first=0) ends at the same solutions
    e2 = _engine(B, prec, active_set_first=0)
    u0b, Xb, Ub, stb, statsb = _run_engine(e2, w, g["fd"])
    assert np.all(stb == 0) and statsb[8:, 1].min() >= 1
    assert rel_err(u0b, g["u0"]) < TOL[prec] and rel_err(Xb, g["X"]) < TOL[prec] and rel_err(Ub, g["U"]) < TOL[prec]


@pytest.mark.parametrize("as_first", [6, 0])
@pytest.mark.parametrize("prec,scale", [("f32", 1.0), ("f64", 1.0), ("f32", 5.0), ("f64", 5.0), ("f32", 15.0), ("f64", 15.0)])
def test_rti_step_vs_oracle(built_lib, c_oracle, prec, scale, as_first):
    """(2) SQP-RTI linearisation + Riccati/IPM QP, config-3 distribution; scale 5 / 15 are the stress
    variants that activate the input bounds."""
    B = 512
    w = wl.independent_problems(B, seed=21, scale=scale)
    fd = np.random.default_rng(4).normal(size=(B, 21, 3))
    e = _engine(B, prec, active_set_first=as_first)
    u0, X, U, st, stats = _run_engine(e, w, fd)
    ou0, oX, oU, ost, r = _run_oracle(c_oracle, make_cfg(), w, fd)
    ok = (ost == 0)
    assert ok.mean() > 0.99
    assert np.all(st[ok] == 0), np.bincount(st[ok])
    tol = TOL[prec] if prec == "f32" else (1e-9 if scale < 10 else 1e-7)
    assert rel_err(u0[ok], ou0[ok]) < tol
    assert rel_err(X[ok], oX[ok]) < tol and rel_err(U[ok], oU[ok]) < tol
    if scale == 1.0:
        assert stats[:, 0].max() == 1 and r["n_active"].max() == 0
    if scale == 15.0:
        assert (r["n_active"] > 0).mean() > 0.5
        if as_first == 0:
            assert stats[r["n_active"] > 0, 1].min() >= 1  # IPM route taken
        print(f"scale 15 {prec} as_first={as_first}: Riccati sweeps mean {stats[:, 0].mean():.2f} max {stats[:, 0].max()}, "
              f"IPM share {(stats[:, 1] > 0).mean():.3f}")


@pytest.mark.parametrize("as_first", [6, 0])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_tightened_bounds(built_lib, c_oracle, prec, as_first):
    """Forced saturation: omega_max = 1.5 rad/s, c_max = 15 m/s^2 (SURVEY.md 8d stress variant)."""
    B = 256
    kw = dict(u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0])
    w = wl.independent_problems(B, seed=31, scale=5.0)
    e = _engine(B, prec, np_=4, active_set_first=as_first, **kw)
    u0, X, U, st, stats = _run_engine(e, w)
    ou0, oX, oU, ost, r = _run_oracle(c_oracle, make_cfg(**kw), w)
    ok = ost == 0
    assert ok.mean() > 0.99 and np.all(st[ok] == 0)
    print(f"tight {prec} as_first={as_first}: Riccati sweeps mean {stats[:, 0].mean():.2f} max {stats[:, 0].max()}, IPM share {(stats[:, 1] > 0).mean():.3f}")
    assert (r["n_active"] > 0).mean() > 0.8
    tol = TOL[prec] if prec == "f32" else 1e-8
    assert rel_err(u0[ok], ou0[ok]) < tol and rel_err(U[ok], oU[ok]) < tol and rel_err(X[ok], oX[ok]) < tol
    assert np.all(U[ok][:, :, :3] <= 1.5 + 1e-6) and np.all(U[ok][:, :, :3] >= -1.5 - 1e-6)


@pytest.mark.parametrize("as_first", [16, 0])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_binding_velocity_boxes(built_lib, c_oracle, prec, as_first):
    """Active, FEASIBLE velocity bounds (lbx/ubx on idx 3,4,5, stages 1..N-1, nmpc_body_rate_ctl.py:59-61,66; VERDICT r1
    item 1a): v_max tightened to ~79 % of the peak speed of the high-dynamics eight, problems whose reference runs
    through the box later in the horizon.  Three warm-started RTI steps on both QP routes; the reference is the C oracle,
    which test_oracle.py::test_binding_velocity_boxes_dense_kkt_vs_c ties to the dense-KKT solve on this workload."""
    B = 128
    vm = np.array(wl.V_BOX)
    kw = dict(v_min=list(-vm), v_max=list(vm))
    w = wl.velocity_box_problems(B, seed=12)
    rng = np.random.default_rng(13)
    x0_seq = [w["x0"] + 0.01 * s * rng.normal(size=w["x0"].shape) * np.array([1, 1, 1, 1, 1, 1, 0, 0, 0, 0]) for s in range(3)]
    e = _engine(B, prec, np_=4, active_set_first=as_first, **kw)
    dt, dev = e.dtype, e.device
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    xr, ur = t(w["xr"]), t(w["ur"])
    e.reset(xr, ur)
    e.set_reference(xr, ur, None)
    X, U = w["xr"].copy(), w["ur"].copy()
    cfg = make_cfg(**kw)
    tol = TOL[prec] if prec == "f32" else 1e-8
    for s in range(3):
        r = c_oracle.rti_batch(cfg, x0_seq[s], w["xr"], w["ur"], None, X, U)
        u0 = e.solve(t(x0_seq[s]))
        torch.cuda.synchronize()
        st, stats = e.status().cpu().numpy(), e.stats().cpu().numpy()
        assert np.all(r["status"] == 0) and np.all(st == 0), (s, np.bincount(st))
        on_box = (np.abs(np.abs(X[:, 1:20, 3:6]) - vm) < 1e-9).sum((1, 2))
        assert (on_box >= 3).mean() > 0.9  # the velocity bound binds over several stages in (almost) every problem
        gX, gU = e.get_all("x").cpu().numpy().astype(np.float64), e.get_all("u").cpu().numpy().astype(np.float64)
        assert rel_err(u0.cpu().numpy(), r["u0"]) < tol, (s, rel_err(u0.cpu().numpy(), r["u0"]))
        assert rel_err(gX, X) < tol and rel_err(gU, U) < tol, (s, rel_err(gX, X), rel_err(gU, U))
        assert np.all(np.abs(gX[:, 1:20, 3:6]) <= vm * (1 + 1e-5))
        print(f"v-box {prec} as_first={as_first} step {s}: u0 err {rel_err(u0.cpu().numpy(), r['u0']):.2e}, Riccati sweeps mean "
              f"{stats[:, 0].mean():.2f} max {stats[:, 0].max()}, IPM share {(stats[:, 1] > 0).mean():.2f}")


def test_failed_problem_keeps_its_iterate(built_lib, c_oracle):
    """a problem whose step is NaN reports status 1, keeps its previous iterate and first input (acados
    SQP_RTI does not update the iterate when the QP fails), does not disturb its neighbours, and solves normally on the
    next call."""
    B = 12
    w = wl.independent_problems(B, seed=91)
    e = _engine(B, "f32", np_=4)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    xr, ur = t(w["xr"]), t(w["ur"])
    e.reset(xr, ur)
    e.set_reference(xr, ur, None)
    x0 = w["x0"].copy()
    x0[5, 2] = np.nan
    u0 = e.solve(t(x0)).cpu().numpy()
    st = e.status().cpu().numpy()
    assert st[5] == 1 and np.all(np.delete(st, 5) == 0)
    X, U = e.get_all("x").cpu().numpy(), e.get_all("u").cpu().numpy()
    assert np.array_equal(X[5], w["xr"][5].astype(np.float32)) and np.array_equal(U[5], w["ur"][5].astype(np.float32))
    assert np.array_equal(u0[5], w["ur"][5, 0].astype(np.float32))
    oX, oU = w["xr"].copy(), w["ur"].copy()
    r = c_oracle.rti_batch(make_cfg(), w["x0"], w["xr"], w["ur"], None, oX, oU)
    keep = np.arange(B) != 5
    assert rel_err(u0[keep], r["u0"][keep]) < 1e-4
    u0b = e.solve(t(w["x0"])).cpu().numpy()  # the poisoned problem recovers from its untouched iterate
    assert np.all(e.status().cpu().numpy() == 0) and rel_err(u0b[5:6], r["u0"][5:6]) < 1e-4


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_warm_started_closed_loop(built_lib, c_oracle, prec):
    """Iterate persists between calls without shifting (nmpc_body_rate_ctl.py:86-112): five RTI steps
    with moving x0 stay on the oracle's trajectory."""
    B = 64
    w = wl.independent_problems(B, seed=41)
    rng = np.random.default_rng(5)
    x0_seq = [w["x0"] + 0.02 * s * rng.normal(size=w["x0"].shape) for s in range(5)]
    e = _engine(B, prec, np_=4)
    u0, X, U, st, _ = _run_engine(e, w, steps=5, x0_seq=x0_seq)
    ou0, oX, oU, ost, _ = _run_oracle(c_oracle, make_cfg(), w, steps=5, x0_seq=x0_seq)
    assert np.all(st == 0) and np.all(ost == 0)
    assert rel_err(u0, ou0) < TOL[prec] and rel_err(X, oX) < TOL[prec]


@pytest.mark.parametrize("N", [40, 80])
def test_longer_horizons(built_lib, c_oracle, N):
    """Config 5 horizons (th_pred stays 0.1 s)."""
    B = 64
    w = wl.independent_problems(B, N=N, seed=51, scale=3.0)
    e = _engine(B, "f32", N=N, np_=4)
    u0, X, U, st, _ = _run_engine(e, w)
    ou0, oX, oU, ost, _ = _run_oracle(c_oracle, make_cfg(N=N), w)
    ok = ost == 0
    assert np.all(st[ok] == 0)
    assert rel_err(u0[ok], ou0[ok]) < 1e-4 and rel_err(X[ok], oX[ok]) < 1e-4


def test_ragged_and_tiny_batches(built_lib, c_oracle):
    """Batch sizes that do not fill a CTA / a half-warp pair, and more problems than resident slots."""
    for B in (1, 3, 9, 6001):
        w = wl.independent_problems(B, seed=B)
        e = _engine(B, "f32", np_=4)
        u0, X, U, st, _ = _run_engine(e, w)
        ou0, oX, oU, ost, _ = _run_oracle(c_oracle, make_cfg(), w)
        assert np.all(st == 0)
        assert rel_err(u0, ou0) < 1e-4 and rel_err(X, oX) < 1e-4


def test_full_size_properties(built_lib):
    """BASELINE config 3 at full size (B = 4096): size-independent properties.
    (i) a problem whose x0 sits on a dynamically consistent iterate = reference returns ~zero step;
    (ii) batch-permutation equivariance; (iii) solving twice from the same state is idempotent in
    the sense that the second RTI step is a much smaller correction."""
    B = 4096
    w = wl.independent_problems(B, seed=0)
    e = _engine(B, "f32", np_=4)
    u0, X, U, st, _ = _run_engine(e, w)
    assert np.all(st == 0)
    perm = np.random.default_rng(0).permutation(B)
    wp = {k: v[perm] for k, v in w.items()}
    e2 = _engine(B, "f32", np_=4)
    u0p, Xp, _, _, _ = _run_engine(e2, wp)
    assert np.array_equal(u0p, u0[perm]) and np.array_equal(Xp, X[perm])
    u0b, Xb, _, stb, _ = _run_engine(e, w, steps=2)
    assert np.abs(Xb - X).max() < 0.25 * np.abs(X - w["xr"]).max()
    assert np.all(np.isfinite(u0)) and np.all(U[:, :, 3] >= 0) and np.all(np.abs(U[:, :, :3]) <= 6 + 1e-5)


def test_infeasible_qp_reports_status(built_lib):
    """Velocity box tighter than the initial speed -> infeasible QP -> non-zero status, and the
    drop-in controller raises the reference's exception (nmpc_body_rate_ctl.py:109-110)."""
    from ndp_nmpc_qd_b200.nmpc_ctl import NMPCBodyRateController

    ctl = NMPCBodyRateController(v_min=[-0.5] * 3, v_max=[0.5] * 3, ipm_max_iter=30)
    xr, ur = wl.reference_horizon([2.75], name="eight_high_dyn")
    ctl.reset(xr[0], ur[0])
    with pytest.raises(Exception, match="acados acados_ocp_solver returned status"):
        ctl.update(xr[0, 0], xr[0], ur[0])


def test_fused_update_equals_set_reference_plus_solve(built_lib):
    """ndp_update (reference upload fused into the solve) is bit-identical to ndp_set_reference +
    ndp_solve, and leaves yref / p stored as if set."""
    for B in (300, 33, 1):  # odd sizes: the last warp of the nominal kernel carries a half without a problem of its own
        w = wl.independent_problems(B, seed=77, scale=5.0)
        fd = np.random.default_rng(1).normal(size=(B, 21, 3))
        e1, e2 = _engine(B, "f32"), _engine(B, "f32")
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
        xr, ur, f, x0 = t(w["xr"]), t(w["ur"]), t(fd), t(w["x0"])
        for e in (e1, e2):
            e.reset(xr, ur)
        e1.set_reference(xr, ur, f)
        u1 = e1.solve(x0)
        u2 = e2.update(x0, xr, ur, f)
        assert torch.equal(u1, u2) and torch.equal(e1.get_all("x"), e2.get_all("x"))
        assert torch.equal(e1.get_all("yref"), e2.get_all("yref")) and torch.equal(e1.get_all("p"), e2.get_all("p"))
        u3 = e2.update(x0, xr, ur, None)  # f = NULL -> zero forces
        e1.set_reference(xr, ur, None)
        assert torch.equal(e1.solve(x0), u3)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_warm_active_set_closed_loop(built_lib, c_oracle, prec):
    """Saturating closed loop (tightened input bounds, moving x0): a solve that ends with active bounds leaves them
    as the first guess of the next solve (active_set_warm).  Every step must still be the oracle's QP solution, the
    cold route (active_set_warm=0) must agree, and the warm route must need fewer Riccati sweeps."""
    B, steps = 128, 6
    kw = dict(u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0])
    w = wl.independent_problems(B, seed=61, scale=5.0)
    rng = np.random.default_rng(8)
    x0_seq = [w["x0"] + 0.03 * s * rng.normal(size=w["x0"].shape) for s in range(steps)]
    ew, ec = _engine(B, prec, np_=4, active_set_warm=1, **kw), _engine(B, prec, np_=4, active_set_warm=0, **kw)
    dt, dev = ew.dtype, ew.device
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    xr, ur = t(w["xr"]), t(w["ur"])
    X, U = w["xr"].copy(), w["ur"].copy()
    cfg = make_cfg(**kw)
    sweeps = {"warm": [], "cold": []}
    for e in (ew, ec):
        e.reset(xr, ur)
        e.set_reference(xr, ur, None)
    # six warm-started steps compound the single-step errors (1e-4 / 1e-8 gates): the iterate of step s is the
    # linearisation point of step s + 1
    # (fp32: one of the 128 problems drifts to 49 of its 80 inputs saturated by step 4; fp64 stays at 5e-13 throughout)
    tol = 5 * TOL[prec] if prec == "f32" else 1e-7
    for s in range(steps):
        r = c_oracle.rti_batch(cfg, x0_seq[s], w["xr"], w["ur"], None, X, U)
        ok = r["status"] == 0
        for name, e in (("warm", ew), ("cold", ec)):
            u0 = e.solve(t(x0_seq[s]))
            torch.cuda.synchronize()
            st, stats = e.status().cpu().numpy(), e.stats().cpu().numpy()
            assert np.all(st[ok] == 0), (name, s, np.bincount(st[ok]))
            assert rel_err(u0.cpu().numpy().astype(np.float64)[ok], r["u0"][ok]) < tol, (name, s)
            assert rel_err(e.get_all("u").cpu().numpy().astype(np.float64)[ok], U[ok]) < tol, (name, s)
            assert rel_err(e.get_all("x").cpu().numpy().astype(np.float64)[ok], X[ok]) < tol, (name, s)
            sweeps[name].append(float(stats[:, 0].mean()))
        assert (r["n_active"] > 0).mean() > 0.3
    print(f"warm active set {prec}: Riccati sweeps per solve warm {np.round(sweeps['warm'], 2)} cold {np.round(sweeps['cold'], 2)}")
    assert np.mean(sweeps["warm"][1:]) < np.mean(sweeps["cold"][1:])
    # reset drops the guess: the first solve after it takes the cold route again
    ew.reset(xr, ur)
    ew.solve(t(x0_seq[0]))
    ec.reset(xr, ur)
    ec.solve(t(x0_seq[0]))
    torch.cuda.synchronize()
    assert np.array_equal(ew.stats().cpu().numpy(), ec.stats().cpu().numpy())
