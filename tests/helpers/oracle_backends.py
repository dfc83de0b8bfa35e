"""Oracle-backed stand-ins with the reference's controller / estimator / plant interfaces (test infrastructure):
the same scenario drivers (ndp_nmpc_qd_b200.formation.ThreeQuadFormation, the config-1 loop) run once on these and
once on the drop-in CUDA classes."""
import numpy as np

from oracle import mlp_numpy
from oracle import refgen_numpy as orf
from oracle.c_oracle import COracle, make_cfg
from oracle.plant_numpy import PlantOracle


class OracleController:
    """NMPCBodyRateController / NDPNMPCBodyRateController interface over oracle/nmpc_oracle.c (fp64)."""

    def __init__(self, c_oracle: COracle, ndp: bool = False, **cfg_kw):
        self.co, self.ndp, self.cfg = c_oracle, ndp, make_cfg(**cfg_kw)
        self.X, self.U = np.zeros((1, 21, 10)), np.zeros((1, 20, 4))
        self.status = 0
        self.n_active = 0

    def reset(self, xr, ur):
        self.X[0], self.U[0] = xr, ur

    def update(self, x0, xr, ur, f=None):
        fd = None if f is None else np.asarray(f, dtype=np.float64)[None]
        r = self.co.rti_batch(self.cfg, np.asarray(x0)[None], np.asarray(xr)[None], np.asarray(ur)[None], fd, self.X, self.U)
        self.status, self.n_active = int(r["status"][0]), int(r["n_active"][0])
        assert self.status == 0, self.status
        return r["u0"][0].copy()


def oracle_downwash_update(weights):
    return lambda other, ego: mlp_numpy.downwash_update(weights, other, ego)


class OraclePlant:
    """MulQuadrotors.forward on [n, 35] numpy states (oracle/plant_numpy.py, pinned to the reference's TorchScript module)."""

    def __init__(self, n, ts_sim=0.01, ts_ctl=0.01, has_downwash=True, has_motor_model=True, has_battery=True):
        self.p = PlantOracle(n, ts_sim, ts_ctl, has_downwash, has_motor_model, has_battery)

    def forward(self, ts_sim, state, cmd):
        return self.p.forward(ts_sim, np.array(state, copy=True), np.asarray(cmd))


def leader_reference(traj):
    return lambda t: orf.horizon(traj, t, 20, 0.1)
