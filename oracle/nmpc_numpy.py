"""ORACLE (test infrastructure only -- never imported by the product path).

Dense-KKT numpy/fp64 restatement of the reference's NMPC hot path: one acados
SQP_RTI step of the quadrotor body-rate OCP.  This is the independent "truth"
the structured C oracle (nmpc_oracle.c) and the CUDA kernels are checked
against.  PARITY UNPINNED at the acados boundary: acados/HPIPM/CasADi are not
vendored in /root/reference and not installed here, and the reference ships no
golden vector for u0 -- this file restates the documented algorithm and is
anchored on the reference's own OCP definition:

  dynamics            ndp_nmpc/scripts/ndp_nmpc_ctl/ndp_nmpc_body_rate_ctl.py:151-162
                      (nmpc_ctl/nmpc_body_rate_ctl.py:147-158 is the f == 0 case)
  cost output y(x,u)  nmpc_body_rate_ctl.py:163-180, weights :48-53
  bounds              nmpc_body_rate_ctl.py:56-61
  solver options      nmpc_body_rate_ctl.py:71-80  (ERK/RK4 1 step per interval,
                      GAUSS_NEWTON, SQP_RTI, PARTIAL_CONDENSING_HPIPM cond_N = N)
  reset / update      nmpc_body_rate_ctl.py:86-112
  constants           params/nmpc_params.py:5-35, params/fhnp_params.py:9-19

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.
"""
from __future__ import annotations

import dataclasses
import numpy as np

NX, NU = 10, 4


@dataclasses.dataclass
class OcpParams:
    """Constants of the OCP (params/nmpc_params.py:9-35, params/fhnp_params.py:9-19)."""

    N: int = 20
    T: float = 2.0
    mass: float = 1.4844
    gravity: float = 9.81
    Q: np.ndarray = dataclasses.field(
        default_factory=lambda: np.array([300.0, 300.0, 400.0, 10.0, 10.0, 10.0, 0.0, 10.0, 10.0, 100.0])
    )
    R: np.ndarray = dataclasses.field(default_factory=lambda: np.array([10.0, 10.0, 10.0, 5.0]))
    u_min: np.ndarray = dataclasses.field(default_factory=lambda: np.array([-6.0, -6.0, -6.0, 0.0]))
    u_max: np.ndarray = dataclasses.field(default_factory=lambda: np.array([6.0, 6.0, 6.0, 9.81 / 0.36]))
    v_min: np.ndarray = dataclasses.field(default_factory=lambda: np.array([-20.0, -20.0, -20.0]))
    v_max: np.ndarray = dataclasses.field(default_factory=lambda: np.array([20.0, 20.0, 20.0]))

    @property
    def h(self) -> float:
        return self.T / self.N


def f_expl(x, u, fd, p: OcpParams):
    """xdot = f(x, u; fd)   (ndp_nmpc_body_rate_ctl.py:151-162)."""
    vx, vy, vz = x[3], x[4], x[5]
    qw, qx, qy, qz = x[6], x[7], x[8], x[9]
    wx, wy, wz, c = u
    return np.array(
        [
            vx,
            vy,
            vz,
            2 * (qx * qz + qw * qy) * c + fd[0] / p.mass,
            2 * (qy * qz - qw * qx) * c + fd[1] / p.mass,
            (1 - 2 * qx**2 - 2 * qy**2) * c - p.gravity + fd[2] / p.mass,
            (-wx * qx - wy * qy - wz * qz) * 0.5,
            (wx * qw + wz * qy - wy * qz) * 0.5,
            (wy * qw - wz * qx + wx * qz) * 0.5,
            (wz * qw + wy * qx - wx * qy) * 0.5,
        ]
    )


def jac_f(x, u):
    """Analytic df/dx (10x10) and df/du (10x4) of f_expl (forces enter additively)."""
    qw, qx, qy, qz = x[6], x[7], x[8], x[9]
    wx, wy, wz, c = u
    A = np.zeros((NX, NX))
    B = np.zeros((NX, NU))
    A[0, 3] = A[1, 4] = A[2, 5] = 1.0
    A[3, 6:10] = 2 * c * np.array([qy, qz, qw, qx])
    A[4, 6:10] = 2 * c * np.array([-qx, -qw, qz, qy])
    A[5, 6:10] = np.array([0.0, -4 * c * qx, -4 * c * qy, 0.0])
    A[6, 6:10] = 0.5 * np.array([0.0, -wx, -wy, -wz])
    A[7, 6:10] = 0.5 * np.array([wx, 0.0, wz, -wy])
    A[8, 6:10] = 0.5 * np.array([wy, -wz, 0.0, wx])
    A[9, 6:10] = 0.5 * np.array([wz, wy, -wx, 0.0])
    B[3, 3] = 2 * (qx * qz + qw * qy)
    B[4, 3] = 2 * (qy * qz - qw * qx)
    B[5, 3] = 1 - 2 * qx**2 - 2 * qy**2
    B[6, 0:3] = 0.5 * np.array([-qx, -qy, -qz])
    B[7, 0:3] = 0.5 * np.array([qw, -qz, qy])
    B[8, 0:3] = 0.5 * np.array([qz, qw, -qx])
    B[9, 0:3] = 0.5 * np.array([-qy, qx, qw])
    return A, B


def rk4_sens(x, u, fd, p: OcpParams):
    """One explicit RK4 step of length h with forward sensitivities.

    The variational equations  Sx' = A_c Sx,  Su' = A_c Su + B_c  are evaluated at
    the RK stage states, i.e. the result is the exact Jacobian of the discrete
    map -- what acados' ERK integrator returns with sens_forw.
    """
    h = p.h

    def vde(xs, Sx, Su):
        A, B = jac_f(xs, u)
        return f_expl(xs, u, fd, p), A @ Sx, A @ Su + B

    Sx0, Su0 = np.eye(NX), np.zeros((NX, NU))
    k1 = vde(x, Sx0, Su0)
    k2 = vde(x + 0.5 * h * k1[0], Sx0 + 0.5 * h * k1[1], Su0 + 0.5 * h * k1[2])
    k3 = vde(x + 0.5 * h * k2[0], Sx0 + 0.5 * h * k2[1], Su0 + 0.5 * h * k2[2])
    k4 = vde(x + h * k3[0], Sx0 + h * k3[1], Su0 + h * k3[2])
    comb = lambda i: (h / 6.0) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i])
    return x + comb(0), Sx0 + comb(1), Su0 + comb(2)


def quat_err_matrix(qr):
    """Rows M1..M3 of qe = M(q_r) q  (nmpc_body_rate_ctl.py:164-166); row 0 is zero."""
    w, x, y, z = qr
    M = np.zeros((4, 4))
    M[1] = [-x, w, -z, y]
    M[2] = [-y, z, w, -x]
    M[3] = [-z, -y, x, w]
    return M


def stage_cost(xk, uk, xrk, urk, p: OcpParams, terminal: bool):
    """Gauss-Newton blocks of the NONLINEAR_LS cost.

    y = [p; v; qwr; qe + q_r(xyz); u], yref = [xr; ur]  =>  residual
    r = [p - pr; v - vr; 0; M q; u - ur]; J is constant.  Stage cost is scaled by
    the shooting interval h (acados cost discretisation), the terminal is not.
    Returns (Hxx[10,10], gx[10], Huu_diag[4], gu[4]).
    """
    s = 1.0 if terminal else p.h
    qr = xrk[6:10]
    M = quat_err_matrix(qr)
    D = np.diag(p.Q[6:10])
    Hxx = np.zeros((NX, NX))
    Hxx[0:6, 0:6] = np.diag(p.Q[0:6])
    Hqq = M.T @ D @ M
    Hxx[6:10, 6:10] = Hqq
    gx = np.zeros(NX)
    gx[0:6] = p.Q[0:6] * (xk[0:6] - xrk[0:6])
    gx[6:10] = Hqq @ xk[6:10]
    if terminal:
        return s * Hxx, s * gx, None, None
    return s * Hxx, s * gx, s * p.R.copy(), s * p.R * (uk - urk)


class DenseQP:
    """min 1/2 z'Hz + g'z  s.t.  Aeq z = beq,  G z <= d   with z = [dx0,du0,dx1,...,dxN]."""

    def __init__(self, N):
        self.N = N
        self.nz = N * (NX + NU) + NX

    def ix(self, k):
        o = k * (NX + NU)
        return slice(o, o + NX)

    def iu(self, k):
        o = k * (NX + NU) + NX
        return slice(o, o + NU)


def build_qp(x0, xr, ur, fd, X, U, p: OcpParams):
    """Linearise the OCP at the iterate (X, U)  (SURVEY.md A.4)."""
    N = p.N
    qp = DenseQP(N)
    nz = qp.nz
    H = np.zeros((nz, nz))
    g = np.zeros(nz)
    Aeq = np.zeros((NX * (N + 1), nz))
    beq = np.zeros(NX * (N + 1))
    # initial condition
    Aeq[0:NX, qp.ix(0)] = np.eye(NX)
    beq[0:NX] = x0 - X[0]
    Grows, drows = [], []
    lin = []
    for k in range(N):
        xn, Sx, Su = rk4_sens(X[k], U[k], fd[k], p)
        bk = xn - X[k + 1]
        lin.append((Sx, Su, bk))
        r = slice(NX * (k + 1), NX * (k + 2))
        Aeq[r, qp.ix(k)] = -Sx
        Aeq[r, qp.iu(k)] = -Su
        Aeq[r, qp.ix(k + 1)] = np.eye(NX)
        beq[r] = bk
        Hxx, gx, Huu, gu = stage_cost(X[k], U[k], xr[k], ur[k], p, False)
        H[qp.ix(k), qp.ix(k)] = Hxx
        g[qp.ix(k)] = gx
        H[qp.iu(k), qp.iu(k)] = np.diag(Huu)
        g[qp.iu(k)] = gu
        # input box, stages 0..N-1
        for i in range(NU):
            e = np.zeros(nz)
            e[qp.iu(k).start + i] = 1.0
            Grows.append(e)
            drows.append(p.u_max[i] - U[k][i])
            Grows.append(-e)
            drows.append(-(p.u_min[i] - U[k][i]))
        # velocity box, stages 1..N-1 (acados lbx/ubx: intermediate nodes only)
        if k >= 1:
            for i in range(3):
                e = np.zeros(nz)
                e[qp.ix(k).start + 3 + i] = 1.0
                Grows.append(e)
                drows.append(p.v_max[i] - X[k][3 + i])
                Grows.append(-e)
                drows.append(-(p.v_min[i] - X[k][3 + i]))
    Hxx, gx, _, _ = stage_cost(X[N], None, xr[N], None, p, True)
    H[qp.ix(N), qp.ix(N)] = Hxx
    g[qp.ix(N)] = gx
    qp.H, qp.g, qp.Aeq, qp.beq = H, g, Aeq, beq
    qp.G, qp.d = np.array(Grows), np.array(drows)
    qp.lin = lin
    return qp


def solve_qp_dense(qp: DenseQP, tol=1e-10, max_iter=60, mu0=10.0, thr=0.1):
    """Mehrotra predictor-corrector primal-dual IPM on the dense KKT system.

    Returns (z, lam, n_iter, status) with status 0 = converged, 2 = max iterations.
    """
    H, g, A, b, G, d = qp.H, qp.g, qp.Aeq, qp.beq, qp.G, qp.d
    nz, ne, ni = H.shape[0], A.shape[0], G.shape[0]
    z = np.zeros(nz)
    pi = np.zeros(ne)
    t = np.maximum(d - G @ z, thr)
    lam = mu0 / t

    def newton(Gam, rhs_z, rhs_e):
        K = np.block([[H + G.T @ (Gam[:, None] * G), A.T], [A, np.zeros((ne, ne))]])
        sol = np.linalg.solve(K, np.concatenate([rhs_z, rhs_e]))
        return sol[:nz], sol[nz:]

    def max_step(v, dv):
        neg = dv < 0
        return min(1.0, np.min(-v[neg] / dv[neg])) if np.any(neg) else 1.0

    status, it = 2, 0
    for it in range(max_iter + 1):
        r_g = H @ z + g + A.T @ pi + G.T @ lam
        r_b = A @ z - b
        r_d = G @ z + t - d
        mu = lam @ t / ni
        if max(np.abs(r_g).max(), np.abs(r_b).max(), np.abs(r_d).max()) < tol and mu < tol:
            status = 0
            break
        if it == max_iter:
            break
        Gam = lam / t

        def direction(r_m):
            dz, dpi = newton(Gam, -(r_g + G.T @ ((lam * r_d - r_m) / t)), -r_b)
            dt = -r_d - G @ dz
            dlam = -(r_m + lam * dt) / t
            return dz, dpi, dt, dlam

        dz, dpi, dt, dlam = direction(lam * t)
        a_aff = min(max_step(lam, dlam), max_step(t, dt))
        mu_aff = (lam + a_aff * dlam) @ (t + a_aff * dt) / ni
        sigma = (mu_aff / mu) ** 3
        dz, dpi, dt, dlam = direction(lam * t + dlam * dt - sigma * mu)
        a = min(1.0, 0.995 * min(max_step(lam, dlam), max_step(t, dt)) if True else 1.0)
        a = min(a, 1.0)
        z = z + a * dz
        pi = pi + a * dpi
        t = t + a * dt
        lam = lam + a * dlam
    return z, lam, it, status


def rti_step(x0, xr, ur, fd, X, U, p: OcpParams | None = None, tol=1e-10):
    """One SQP_RTI step (linearise at (X,U), solve the QP, full step).

    x0[10], xr[N+1,10], ur[N,4], fd[N+1,3] (Newtons; zeros for the plain NMPC
    controller), X[N+1,10], U[N,4] the persistent iterate (no shift between calls,
    nmpc_body_rate_ctl.py:86-112).  Returns dict(u0, X, U, status, n_iter, n_active).
    """
    p = p or OcpParams()
    qp = build_qp(x0, xr, ur, fd, X, U, p)
    z, lam, n_iter, qp_status = solve_qp_dense(qp, tol=tol)
    Xn, Un = X.copy(), U.copy()
    for k in range(p.N):
        Xn[k] = X[k] + z[qp.ix(k)]
        Un[k] = U[k] + z[qp.iu(k)]
    Xn[p.N] = X[p.N] + z[qp.ix(p.N)]
    status = 0 if qp_status == 0 else 4  # acados: QP failure -> status 4
    if not np.all(np.isfinite(z)):
        status = 1
    slack = qp.d - qp.G @ z
    n_active = int(np.sum((slack < 1e-6) & (lam > 1e-6)))
    return dict(u0=Un[0].copy(), X=Xn, U=Un, status=status, n_iter=n_iter, n_active=n_active, qp=qp, z=z)
