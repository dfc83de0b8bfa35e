"""CPU tests of the oracle itself: pinned against the golden vectors generated from the
reference's own code (tests/golden/make_golden.py), the SURVEY.md known answers and an independent
dense-KKT solve."""
import numpy as np
import pytest

from conftest import golden
from oracle import mlp_numpy, nmpc_numpy as on
from oracle.c_oracle import make_cfg
from ndp_nmpc_qd_b200 import workloads as wl

# SURVEY.md appendix B.5 (fp64 known answers of RK4 + sensitivities)
X_B5 = np.array([0.1, -0.2, 0.3, 1.0, -2.0, 0.5, 0.9, 0.1, -0.2, 0.3])
U_B5 = np.array([1.2, -0.7, 0.4, 13.0])
F_B5 = np.array([0.3, -0.2, -4.0])
XN_B5 = np.array([0.182029795955286, -0.422847260908461, 0.345105206525358, 0.646352242513726, -2.483264040529536,
                  0.392632435728246, 0.878666319606771, 0.160186178335938, -0.214964229588542, 0.325693514174479])
SX3_B5 = np.array([0, 0, 0, 1, 0, 0, -0.551135048137567, 0.80000544543836, 2.299034610188333, 0.427040113258906])
SU3_B5 = np.array([-0.001460054872415, -0.00170566061877, 0.004432974139577, -0.028758303339999, -0.036137737172338,
                   0.087930890885466, 0, 0, 0, 0])


def test_dynamics_known_answer(c_oracle):
    f_ref = np.array([1, -2, 0.5, -3.697898140662895, -4.034734572891404, -0.80469145782808, -0.19, 0.605, -0.155, 0.265])
    assert np.allclose(on.f_expl(X_B5, U_B5, F_B5, on.OcpParams()), f_ref, atol=1e-13)
    assert np.allclose(c_oracle.f(make_cfg(), X_B5, U_B5, F_B5), f_ref, atol=1e-13)


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_rk4_sens_known_answer(c_oracle, impl):
    if impl == "numpy":
        xn, Sx, Su = on.rk4_sens(X_B5, U_B5, F_B5, on.OcpParams())
    else:
        xn, Sx, Su = c_oracle.rk4_sens(make_cfg(), X_B5, U_B5, F_B5)
    assert np.allclose(xn, XN_B5, atol=1e-13)
    assert np.allclose(Sx[3], SX3_B5, atol=1e-12)
    assert np.allclose(Su[:, 3], SU3_B5, atol=1e-12)
    assert np.allclose(Su[3, 0:3], [0.039758937219711, 0.043378980900077, 0.00265996720885], atol=1e-12)


def test_sensitivities_vs_central_differences(c_oracle):
    cfg = make_cfg()
    rng = np.random.default_rng(0)
    for _ in range(5):
        x = rng.normal(size=10)
        u = np.array([*rng.normal(size=3), 9.81 + rng.normal()])
        f = rng.normal(size=3)
        _, Sx, Su = c_oracle.rk4_sens(cfg, x, u, f)
        eps = 1e-6
        for j in range(10):
            d = np.zeros(10); d[j] = eps
            col = (c_oracle.rk4_sens(cfg, x + d, u, f)[0] - c_oracle.rk4_sens(cfg, x - d, u, f)[0]) / (2 * eps)
            assert np.allclose(col, Sx[:, j], atol=2e-9)
        for j in range(4):
            d = np.zeros(4); d[j] = eps
            col = (c_oracle.rk4_sens(cfg, x, u + d, f)[0] - c_oracle.rk4_sens(cfg, x, u - d, f)[0]) / (2 * eps)
            assert np.allclose(col, Su[:, j], atol=2e-9)


def test_c_oracle_matches_dense_kkt_golden(c_oracle):
    """Structured Riccati-IPM (C) vs the dense-KKT numpy solutions stored in rti_golden.npz
    (4 nominal, 4 at 5x perturbation, 4 at 15x with up to 10 active bounds)."""
    g = golden("rti_golden.npz")
    X, U = g["xr"].copy(), g["ur"].copy()
    r = c_oracle.rti_batch(make_cfg(), g["x0"], g["xr"], g["ur"], g["fd"], X, U)
    assert np.all(r["status"] == 0)
    assert np.abs(r["u0"] - g["u0"]).max() < 1e-8
    assert np.abs(X - g["X"]).max() < 1e-8 and np.abs(U - g["U"]).max() < 1e-8
    assert np.array_equal(r["n_active"], g["n_active"])
    assert g["n_active"].max() >= 9  # the inequality path is exercised


def test_c_oracle_matches_dense_kkt_live(c_oracle):
    w = wl.independent_problems(2, seed=123, scale=12.0)
    p = on.OcpParams()
    X, U = w["xr"].copy(), w["ur"].copy()
    r = c_oracle.rti_batch(make_cfg(), w["x0"], w["xr"], w["ur"], None, X, U)
    for b in range(2):
        d = on.rti_step(w["x0"][b], w["xr"][b], w["ur"][b], np.zeros((21, 3)), w["xr"][b].copy(), w["ur"][b].copy(), p)
        assert d["status"] == 0 and r["status"][b] == 0
        assert np.abs(d["u0"] - r["u0"][b]).max() < 1e-8
        assert np.abs(d["X"] - X[b]).max() < 1e-8


def test_binding_velocity_boxes_dense_kkt_vs_c(c_oracle):
    """Feasible problems with ACTIVE velocity bounds (lbx/ubx on idx 3,4,5, stages 1..N-1, nmpc_body_rate_ctl.py:59-61,66;
    VERDICT r1 item 1a): the structured C solve (IPM + exact active-set rounds with pinned velocity components) against
    the independent dense-KKT numpy solve, incl. a second, warm-started RTI step."""
    vm = np.array(wl.V_BOX)
    w = wl.velocity_box_problems(6, seed=2)
    cfg = make_cfg(v_min=list(-vm), v_max=list(vm))
    p = on.OcpParams(v_min=-vm, v_max=vm)
    X, U = w["xr"].copy(), w["ur"].copy()
    Xd, Ud = w["xr"].copy(), w["ur"].copy()
    for step in range(2):
        r = c_oracle.rti_batch(cfg, w["x0"], w["xr"], w["ur"], None, X, U)
        assert np.all(r["status"] == 0)
        for b in range(6):
            d = on.rti_step(w["x0"][b], w["xr"][b], w["ur"][b], np.zeros((21, 3)), Xd[b], Ud[b], p, tol=1e-12)
            Xd[b], Ud[b] = d["X"], d["U"]
            assert d["status"] == 0
            assert np.abs(d["u0"] - r["u0"][b]).max() < 1e-8 and np.abs(d["X"] - X[b]).max() < 1e-8 and np.abs(d["U"] - U[b]).max() < 1e-7
            assert d["n_active"] == r["n_active"][b]
        # a velocity bound is binding (not merely an input bound): the solution sits on the box at several stages
        on_box = (np.abs(np.abs(X[:, 1:20, 3:6]) - vm) < 1e-9).sum((1, 2))
        assert on_box.min() >= 3, on_box
        assert np.all(np.abs(X[:, 1:20, 3:6]) <= vm + 1e-9)


def test_plain_ipm_of_the_c_oracle_is_the_hpipm_like_baseline(c_oracle):
    """make_cfg(tol=1e-8, max_iter=50) -- the amount of work bench.py times as the CPU arm -- stops at the interior-point
    iterate (no polish): close to, but not on, the exact solution when bounds are active."""
    w = wl.independent_problems(16, seed=9, scale=12.0)
    Xa, Ua, Xb, Ub = w["xr"].copy(), w["ur"].copy(), w["xr"].copy(), w["ur"].copy()
    ra = c_oracle.rti_batch(make_cfg(), w["x0"], w["xr"], w["ur"], None, Xa, Ua)
    rb = c_oracle.rti_batch(make_cfg(tol=1e-8, max_iter=50), w["x0"], w["xr"], w["ur"], None, Xb, Ub)
    ok = (ra["status"] == 0) & (rb["status"] == 0)
    assert ok.sum() >= 15 and (ra["n_active"][ok] > 0).any()
    assert np.abs(ra["u0"][ok] - rb["u0"][ok]).max() < 1e-5


def test_qp_kkt_residuals():
    """The dense solve satisfies the KKT conditions of the QP it was given."""
    w = wl.independent_problems(1, seed=5, scale=15.0)
    d = on.rti_step(w["x0"][0], w["xr"][0], w["ur"][0], np.zeros((21, 3)), w["xr"][0].copy(), w["ur"][0].copy())
    qp, z = d["qp"], d["z"]
    assert np.abs(qp.Aeq @ z - qp.beq).max() < 1e-9
    assert (qp.G @ z - qp.d).max() < 1e-8


def test_qp_solution_vs_scipy_slsqp():
    """Third-party pin of the QP solve: the same linearised QP, condensed onto the null space of the dynamics equalities,
    handed to scipy's SLSQP (Kraft's active-set SQP; scipy is an independent, published solver present in this image --
    acados / HPIPM are not).  The QP is strictly convex, so every correct solver, HPIPM included, has ONE answer; SLSQP
    stops at ~1e-7 of it (its line search runs out of precision first: status 8 on the velocity-box cases).  Covers no
    active bound, active input bounds and 14-15 active velocity bounds."""
    so = pytest.importorskip("scipy.optimize")
    import scipy.linalg as sl
    vm = np.array(wl.V_BOX)
    cases = [(wl.independent_problems(2, seed=5), on.OcpParams()), (wl.independent_problems(2, seed=5, scale=15.0), on.OcpParams()),
             (wl.velocity_box_problems(2, seed=2), on.OcpParams(v_min=-vm, v_max=vm))]
    n_act = []
    for w, p in cases:
        for b in range(2):
            d = on.rti_step(w["x0"][b], w["xr"][b], w["ur"][b], np.zeros((21, 3)), w["xr"][b].copy(), w["ur"][b].copy(), p, tol=1e-12)
            qp, z = d["qp"], d["z"]
            Z = sl.null_space(qp.Aeq)
            zp = np.linalg.lstsq(qp.Aeq, qp.beq, rcond=None)[0]
            Hr, gr, Gr, dr = Z.T @ qp.H @ Z, Z.T @ (qp.H @ zp + qp.g), qp.G @ Z, qp.d - qp.G @ zp
            fin = np.isfinite(dr)  # the default velocity box is +-inf
            r = so.minimize(lambda y: 0.5 * y @ Hr @ y + gr @ y, np.zeros(Z.shape[1]), jac=lambda y: Hr @ y + gr, method="SLSQP",
                            constraints=[{"type": "ineq", "fun": lambda y: dr[fin] - Gr[fin] @ y, "jac": lambda y: -Gr[fin]}],
                            options=dict(ftol=1e-16, maxiter=500))
            assert r.status in (0, 8), r.message
            assert np.abs(zp + Z @ r.x - z).max() < 2e-6
            n_act.append(d["n_active"])
    assert n_act[0] == 0 and max(n_act[2:4]) >= 1 and min(n_act[4:]) >= 10, n_act


def test_rti_is_exact_for_feasible_unconstrained_case(c_oracle):
    """No active bound => the RTI step equals the equality-constrained LQ solution: the IPM path
    and a single Riccati sweep must agree (basis of the CUDA fast path)."""
    w = wl.independent_problems(8, seed=3)
    X, U = w["xr"].copy(), w["ur"].copy()
    r = c_oracle.rti_batch(make_cfg(), w["x0"], w["xr"], w["ur"], None, X, U)
    big = make_cfg(u_min=[-1e6] * 4, u_max=[1e6] * 4, v_min=[-1e6] * 3, v_max=[1e6] * 3)
    X2, U2 = w["xr"].copy(), w["ur"].copy()
    r2 = c_oracle.rti_batch(big, w["x0"], w["xr"], w["ur"], None, X2, U2)
    assert r["n_active"].max() == 0
    assert np.abs(r["u0"] - r2["u0"]).max() < 1e-8 and np.abs(X - X2).max() < 1e-8


def test_horizon_40_and_warm_start(c_oracle):
    for N in (40, 80):
        cfg = make_cfg(N=N)
        w = wl.independent_problems(4, N=N, seed=9)
        X, U = w["xr"].copy(), w["ur"].copy()
        r = c_oracle.rti_batch(cfg, w["x0"], w["xr"], w["ur"], None, X, U)
        assert np.all(r["status"] == 0)
        # second RTI step from the warm start changes the iterate much less than the first
        X1 = X.copy()
        c_oracle.rti_batch(cfg, w["x0"], w["xr"], w["ur"], None, X, U)
        assert np.abs(X - X1).max() < 0.2 * max(np.abs(X1 - w["xr"]).max(), 1e-3)


def test_mlp_oracle_vs_reference_module(mlp_weights):
    """numpy MLP vs the reference's torch nn.Sequential with the shipped weights (golden)."""
    g = golden("mlp_golden.npz")
    y32 = mlp_numpy.mlp_forward(mlp_weights, g["x"], np.float32)
    y64 = mlp_numpy.mlp_forward(mlp_weights, g["x"], np.float64)
    assert np.abs(y64 - g["y64"]).max() < 1e-12
    assert np.abs(y32 - g["y32"]).max() < 2e-5
    # SURVEY.md B.1 first row
    assert np.allclose(g["y32"][0], [0.11374213, -0.85416198, -4.45603561], atol=2e-6)


def test_downwash_update_oracle(mlp_weights):
    g = golden("downwash_golden.npz")
    f = mlp_numpy.downwash_update(mlp_weights, g["other"], g["ego"])
    assert f.dtype == np.float32 and f.shape == (21, 3)
    assert np.abs(f - g["f"]).max() < 2e-5
