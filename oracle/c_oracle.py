"""ctypes binding of oracle/nmpc_oracle.c (ORACLE -- test infrastructure only).

Mirrors the argument layout of the reference controller's update()
(ndp_nmpc/scripts/ndp_nmpc_ctl/ndp_nmpc_body_rate_ctl.py:91-112), batched over
problems: x0[B,10], xr[B,N+1,10], ur[B,N,4], f[B,N+1,3], iterate X[B,N+1,10], U[B,N,4].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NX, NU = 10, 4


class OrcCfg(C.Structure):
    _fields_ = [
        ("N", C.c_int),
        ("h", C.c_double),
        ("mass", C.c_double),
        ("gravity", C.c_double),
        ("Q", C.c_double * 10),
        ("R", C.c_double * 4),
        ("u_min", C.c_double * 4),
        ("u_max", C.c_double * 4),
        ("v_min", C.c_double * 3),
        ("v_max", C.c_double * 3),
        ("tol", C.c_double),
        ("tol_mu", C.c_double),
        ("max_iter", C.c_int),
        ("mu0", C.c_double),
        ("t_floor", C.c_double),
        ("polish", C.c_int),
        ("pdas_first", C.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Returns the fp64 library path."""
    out = os.path.join(_HERE, "_build", "libnmpc_oracle.so")
    src = os.path.join(_HERE, "nmpc_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)
    return out


def make_cfg(N=20, T=None, tol=None, tol_mu=None, max_iter=100, u_min=None, u_max=None, v_min=None, v_max=None, polish=None, pdas_first=0) -> OrcCfg:
    """Constants of params/nmpc_params.py:9-35 and params/fhnp_params.py:9-19.

    Default (tol=None): interior-point iterations to 1e-9 -- far enough to identify the active set while the
    barrier-weighted Riccati recursion is still accurate -- followed by exact primal-dual active-set rounds
    (nmpc_oracle.c: orc_polish), which land on the KKT point of the QP itself (no barrier floor; needed for active
    velocity bounds, where the plain IPM of this file is only good to ~3e-5).  An explicit tol gives the plain IPM
    without the polish: make_cfg(tol=1e-8, max_iter=50) is the HPIPM-like amount of work the CPU baseline times.
    th_pred = T/N is 0.1 s in the reference; for N != 20 the horizon is T = 0.1 N
    (th_pred must stay a multiple of ts_nmpc).
    """
    c = OrcCfg()
    c.N = N
    c.h = (T if T is not None else 0.1 * N) / N
    c.mass, c.gravity = 1.4844, 9.81
    c.Q[:] = [300, 300, 400, 10, 10, 10, 0, 10, 10, 100]
    c.R[:] = [10, 10, 10, 5]
    c.u_min[:] = list(u_min) if u_min is not None else [-6, -6, -6, 0]
    c.u_max[:] = list(u_max) if u_max is not None else [6, 6, 6, 9.81 / 0.36]
    c.v_min[:] = list(v_min) if v_min is not None else [-20, -20, -20]
    c.v_max[:] = list(v_max) if v_max is not None else [20, 20, 20]
    # default: IPM to 1e-9 (enough to identify the active set while the barrier-weighted Riccati recursion is still
    # accurate) + exact active-set polish; an explicit tol gives the plain IPM (HPIPM-like work: tol=1e-8, max_iter=50)
    if polish is None:
        polish = 150 if tol is None else 0  # rounds; the damped phase after a detected cycle releases one bound per round
    if tol is None:
        tol = 1e-9
    c.tol, c.max_iter, c.mu0, c.t_floor = tol, max_iter, 10.0, 0.1
    c.tol_mu = tol if tol_mu is None else tol_mu
    c.polish = int(polish)
    c.pdas_first = int(pdas_first)
    return c


class COracle:
    def __init__(self, precision: str = "f64"):
        path = build()
        if precision == "f32":
            path = path.replace("libnmpc_oracle.so", "libnmpc_oracle_f32.so")
        self.lib = C.CDLL(path)
        self.dtype = np.float64 if self.lib.orc_sizeof_real() == 8 else np.float32
        self.lib.orc_rti_batch.restype = C.c_int

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def rk4_sens(self, cfg: OrcCfg, x, u, fd):
        dt = self.dtype
        x, u, fd = (np.ascontiguousarray(v, dtype=dt) for v in (x, u, fd))
        xn, AB = np.zeros(NX, dt), np.zeros((NX, NX + NU), dt)
        self.lib.orc_rk4_sens(C.byref(cfg), self._p(x), self._p(u), self._p(fd), self._p(xn), self._p(AB))
        return xn, AB[:, :NX].copy(), AB[:, NX:].copy()

    def f(self, cfg, x, u, fd):
        dt = self.dtype
        x, u, fd = (np.ascontiguousarray(v, dtype=dt) for v in (x, u, fd))
        xd = np.zeros(NX, dt)
        self.lib.orc_f(C.byref(cfg), self._p(x), self._p(u), self._p(fd), self._p(xd))
        return xd

    def rti_batch(self, cfg: OrcCfg, x0, xr, ur, fd, X, U, nthreads: int = 0):
        """In-place RTI step on the iterate (X, U).  Returns dict(u0, status, n_iter, n_active, threads)."""
        dt = self.dtype
        N = cfg.N
        x0 = np.ascontiguousarray(x0, dtype=dt).reshape(-1, NX)
        B = x0.shape[0]
        xr = np.ascontiguousarray(xr, dtype=dt).reshape(B, N + 1, NX)
        ur = np.ascontiguousarray(ur, dtype=dt).reshape(B, N, NU)
        fdp = None
        if fd is not None:
            fd = np.ascontiguousarray(fd, dtype=dt).reshape(B, N + 1, 3)
            fdp = self._p(fd)
        assert X.dtype == dt and U.dtype == dt and X.flags.c_contiguous and U.flags.c_contiguous
        assert X.shape == (B, N + 1, NX) and U.shape == (B, N, NU)
        u0 = np.zeros((B, NU), dt)
        status = np.zeros(B, np.int32)
        n_iter = np.zeros(B, np.int32)
        n_active = np.zeros(B, np.int32)
        used = self.lib.orc_rti_batch(
            C.byref(cfg), B, self._p(x0), self._p(xr), self._p(ur), fdp, self._p(X), self._p(U),
            self._p(u0), self._p(status), self._p(n_iter), self._p(n_active), int(nthreads),
        )
        return dict(u0=u0, status=status, n_iter=n_iter, n_active=n_active, threads=used)
