"""NDPNMPCBodyRateController: the body-rate NMPC with three disturbance-force parameters.

Reference: ndp_nmpc/scripts/ndp_nmpc_ctl/ndp_nmpc_body_rate_ctl.py:20-112; p_i = [quaternion_r_i;
f_i] (:97-104), forces enter the velocity dynamics as f / mass (:155-157).
"""
from __future__ import annotations

from ..nmpc_ctl.nmpc_body_rate_ctl import BodyRateControllerBase


class NDPNMPCBodyRateController(BodyRateControllerBase):
    N_PARAMS = 7  # p = [quaternion_r; disturb_f]  (ndp_nmpc_body_rate_ctl.py:197)

    def update(self, x0, xr, ur, f):
        self.solver.set_reference(xr, ur, f)
        u0 = self.solver.solve_for_x0(x0)  # feedback, take the first action
        self._raise_on_status()
        return u0
