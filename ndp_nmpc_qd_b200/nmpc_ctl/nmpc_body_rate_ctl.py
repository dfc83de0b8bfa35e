"""NMPCBodyRateController with the reference's interface, backed by the CUDA engine.

Reference: ndp_nmpc/scripts/nmpc_ctl/nmpc_body_rate_ctl.py:20-112 -- same constructor
argument, `reset(xr, ur)`, `update(x0, xr, ur) -> u0`, `.solver` with the acados surface, and
the same exception when the solver reports a non-zero status.  No acados / CasADi / code
generation is involved: `is_build_acados` is accepted and ignored (the CUDA library is prebuilt).
"""
from __future__ import annotations

import numpy as np

from ..params import nmpc_params as CP
from ..solver import BatchedOcpSolver


class BodyRateControllerBase(object):
    """What the two reference controllers share (construction, reset, status check).  The reference's classes are
    siblings, not parent and child, and nmpc_node.py:203-208 dispatches on them with isinstance -- so they are
    siblings here too (an NDP controller must not pass `isinstance(ctl, NMPCBodyRateController)`)."""

    N_PARAMS = 4

    def __init__(self, is_build_acados=True, batch: int = 1, precision: str = "f32", device="cuda:0", N: int = CP.N_node,
                 **solver_overrides):
        del is_build_acados
        self.batch = batch
        self.solver = BatchedOcpSolver(batch=batch, N=N, np_=self.N_PARAMS, precision=precision, device=device,
                                       **solver_overrides)

    def reset(self, xr, ur):
        # reset x and u of the controller, which prevents warm-starting from the previous solution
        # (nmpc_body_rate_ctl.py:86-91)
        if self.batch == 1:
            for i in range(self.solver.N):
                self.solver.set(i, "x", xr[i, :])
                self.solver.set(i, "u", ur[i, :])
            self.solver.set(self.solver.N, "x", xr[self.solver.N, :])
        else:
            self.solver.reset(xr, ur)

    def _raise_on_status(self):
        st = np.atleast_1d(self.solver.status)
        if np.any(st != 0):
            bad = int(st[np.nonzero(st)[0][0]])
            raise Exception("acados acados_ocp_solver returned status {}. Exiting.".format(bad))


class NMPCBodyRateController(BodyRateControllerBase):
    N_PARAMS = 4  # p = quaternion_r (nmpc_body_rate_ctl.py:193)

    def update(self, x0, xr, ur):
        # yref_i = [xr_i; ur_i], p_i = quaternion_r_i  (nmpc_body_rate_ctl.py:93-104)
        self.solver.set_reference(xr, ur)
        u0 = self.solver.solve_for_x0(x0)  # feedback, take the first action
        self._raise_on_status()
        return u0
