"""CPU oracle (test infrastructure only) of the dop_sim batched quadrotor plant -- SURVEY.md section 8f-1.

numpy float64 restatement of /root/reference/dop_sim/scripts/quadrotor/ (file:line below are relative to
that directory).  Pinned: tests/golden/plant_golden.npz holds outputs of the reference's own TorchScript
module (tests/golden/make_plant_golden.py) and tests/test_plant_oracle.py checks this file against them,
including the SURVEY.md B.4 known answer.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import numpy as np

# ---- params/physical_param.py:34-95 ----
L_FRAME, ALPHA_FRAME = 0.1372, 45.0 * np.pi / 180.0
MASS, GRAVITY = 1.4844, 9.81
IXX, IYY, IZZ, IXZ = 0.0094, 0.0134, 0.0145, 0.0
O_MAX, O_MIN = 24000 / 1000, 2600 / 1000
K_Q, K_T = 3.7611e-10 * 1e6, 2.8158e-08 * 1e6
TM = 0.0840
KD_X, KD_Y, KD_Z, K_H = 0.26, 0.28, 0.42, 0.01
DW_RANGE_HORIZ, DW_RANGE_VERT = 1.5, 4
RP, K_D1, K_D2, K_D3 = 0.0775, 4000, 0.65, -0.10
_LS, _LC = L_FRAME * np.sin(ALPHA_FRAME), L_FRAME * np.cos(ALPHA_FRAME)
G_1 = np.array([[1, 1, 1, 1], [-_LS, _LS, _LS, -_LS], [-_LC, _LC, -_LC, _LC], [-K_Q / K_T, -K_Q / K_T, K_Q / K_T, K_Q / K_T]])
_GAM = IXX * IZZ - IXZ**2
GAMMA1 = (IXZ * (IXX - IYY + IZZ)) / _GAM
GAMMA2 = (IZZ * (IZZ - IYY) + IXZ**2) / _GAM
GAMMA3, GAMMA4 = IZZ / _GAM, IXZ / _GAM
GAMMA5, GAMMA6 = (IZZ - IXX) / IYY, IXZ / IYY
GAMMA7 = ((IXX - IYY) * IXX + IXZ**2) / _GAM
GAMMA8 = IXX / _GAM
# ---- params/control_param.py:17-57 ----
SIGMA, K_TH, B_TH, T_ALL = 0.05, 17.666, -1.206, 705.0
RATE_KP, RATE_KI, RATE_KD = (0.3, 0.3, 0.13), (0.01, 0.01, 0.01), (0.005, 0.005, 0.005)
_CSC, _SEC = 1 / (4 * L_FRAME * np.sin(ALPHA_FRAME)), 1 / (4 * L_FRAME * np.cos(ALPHA_FRAME))
G_1_INV = np.array([[0.25, -_CSC, -_SEC, -K_T / (4 * K_Q)], [0.25, _CSC, _SEC, -K_T / (4 * K_Q)],
                    [0.25, _CSC, -_SEC, K_T / (4 * K_Q)], [0.25, -_CSC, _SEC, K_T / (4 * K_Q)]])
O_MIN_F32, O_MAX_F32 = float(np.float32(O_MIN)), float(np.float32(O_MAX))
HALF_PI_F32 = float(np.float32(math.pi)) / 2  # qd_dynamics.py:36,135: bool tensor * python float is float32 in torch


def _rigid_body(x, u):
    """a_dynamics/rigid_body_use_vw.py:32-108: x[n,13] = (e,n,u, v3, ew ex ey ez, p q r), u[n,6] = (f_i, l m n)."""
    vx, vy, vz, ew, ex, ey, ez, p, q, r = (x[:, i] for i in range(3, 13))
    fx, fy, fz, l, m, n = (u[:, i] for i in range(6))
    return np.stack([
        vx, vy, vz, fx / MASS, fy / MASS, fz / MASS - GRAVITY,
        (0 - p * ex - q * ey - r * ez) / 2, (p * ew + 0 + r * ey - q * ez) / 2,
        (q * ew - r * ex + 0 + p * ez) / 2, (r * ew + q * ex - p * ey + 0) / 2,
        GAMMA1 * p * q - GAMMA2 * q * r + GAMMA3 * l + GAMMA4 * n,
        GAMMA5 * p * r - GAMMA6 * (p**2 - r**2) + m / IYY,
        GAMMA7 * p * q - GAMMA1 * q * r + GAMMA4 * l + GAMMA8 * n,
    ], 1)


def _rotation(q):
    """tools/rotations.py:95-117 (body -> inertial)."""
    e0, e1, e2, e3 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = e1**2 + e0**2 - e2**2 - e3**2; R[:, 0, 1] = 2 * (e1 * e2 - e3 * e0); R[:, 0, 2] = 2 * (e1 * e3 + e2 * e0)
    R[:, 1, 0] = 2 * (e1 * e2 + e3 * e0); R[:, 1, 1] = e2**2 + e0**2 - e1**2 - e3**2; R[:, 1, 2] = 2 * (e2 * e3 - e1 * e0)
    R[:, 2, 0] = 2 * (e1 * e3 - e2 * e0); R[:, 2, 1] = 2 * (e2 * e3 + e1 * e0); R[:, 2, 2] = e3**2 + e0**2 - e1**2 - e2**2
    return R


class PlantOracle:
    """MulQuadrotors (mul_quadrotors.py:19-50): state [n,35] float64, cmd [n,4] (rates rad/s, throttle 0..1)."""

    def __init__(self, num_agent, ts_sim, ts_control, has_downwash=True, has_motor_model=True, has_battery=True):
        self.n, self.ts_ctl = num_agent, ts_control
        self.has_downwash, self.has_motor_model, self.has_battery = has_downwash, has_motor_model, has_battery
        self.ctl_t, self.all_sim_t = 999.0, 0.0
        self.delta = np.zeros((num_agent, 4))
        self.motor_alpha = math.exp(-ts_sim / TM)  # qd_dynamics.py:73
        tsf = ts_control / 0.02                    # b_autopilot/atp_rate.py:22
        self.kp = np.array(RATE_KP); self.ki = np.array(RATE_KI) / tsf; self.kd = np.array(RATE_KD) * tsf
        self.a1 = (2.0 * SIGMA - ts_control) / (2.0 * SIGMA + ts_control)  # pid_control.py:21-24
        self.a2 = 2.0 / (2.0 * SIGMA + ts_control)
        self.integ = np.zeros((num_agent, 3)); self.e_d1 = np.zeros((num_agent, 3)); self.ed_d1 = np.zeros((num_agent, 3))

    def autopilot(self, s, cmd, all_sim_t):
        """b_autopilot/atp_rate.py:60-112 + pid_control.py:36-65."""
        voltage_cf = (4.2 - all_sim_t / T_ALL * (4.2 - 3.6)) / 4.2 if self.has_battery else 1.0
        thrust = 4 * (cmd[:, 3] * K_TH + B_TH) * voltage_cf
        thrust = np.where(thrust < 0, 0.0, thrust)
        err = cmd[:, 0:3] - s[:, 19:22]
        self.integ = self.integ + (self.ts_ctl / 2) * (err + self.e_d1)
        err_dot = self.a1 * self.ed_d1 + self.a2 * (err - self.e_d1)
        u = self.kp * err + self.ki * self.integ + self.kd * err_dot
        u_sat = np.clip(u, -999.0, 999.0)
        self.integ = self.integ + (self.ts_ctl / self.ki) * (u_sat - u)  # |ki| > 1e-4 always here
        self.e_d1, self.ed_d1 = err, err_dot
        tpm = np.concatenate([thrust[:, None], u_sat], 1) @ G_1_INV.T
        tpm = np.where(tpm < 0, 0.0, tpm)
        return np.sqrt(tpm / K_T)

    def dynamics(self, dt, s, delta):
        """a_dynamics/qd_dynamics.py:75-248; s is updated in place and returned."""
        n = s.shape[0]
        if self.has_motor_model:  # :224-228
            delta = self.motor_alpha * s[:, 31:35] + (1 - self.motor_alpha) * delta
        # :230-248; tools/saturate.py:24-30 multiplies bool tensors by python floats, so the limits are float32 values
        d = np.where(delta <= O_MIN, O_MIN_F32, np.where(delta >= O_MAX, O_MAX_F32, delta))
        tt = (K_T * d**2) @ G_1.T
        R = _rotation(s[:, 9:13])
        vwb = np.einsum("nji,nj->ni", R, s[:, 28:31])  # R^T V_wind
        ur, vr, wr = s[:, 16] - vwb[:, 0], s[:, 17] - vwb[:, 1], s[:, 18] - vwb[:, 2]
        Va = np.sqrt(ur**2 + vr**2 + wr**2)
        with np.errstate(invalid="ignore", divide="ignore"):
            s[:, 22] = Va
            s[:, 24] = np.where(ur == 0, HALF_PI_F32, 0.0) + np.where(ur != 0, np.arctan2(wr, ur), 0.0)
            s[:, 25] = np.where(Va != 0, 1.0, 0.0) * np.arcsin(vr / Va)  # NaN when Va == 0 (0 * NaN), as in torch
        fb = np.stack([-KD_X * ur, -KD_Y * vr, tt[:, 0] + (-KD_Z * wr + K_H * (ur**2 + vr**2))], 1)
        f_i = np.einsum("nij,nj->ni", R, fb)
        if self.has_downwash:  # :161-198 (row = ego, column = other)
            dx = s[None, :, 3] - s[:, None, 3]; dy = s[None, :, 4] - s[:, None, 4]; dz = s[None, :, 5] - s[:, None, 5]
            dx = np.where(dx > DW_RANGE_HORIZ, DW_RANGE_HORIZ, dx); dy = np.where(dy > DW_RANGE_HORIZ, DW_RANGE_HORIZ, dy)
            dh = np.sqrt(dx**2 + dy**2)
            valid = (dh < DW_RANGE_HORIZ) & (dz > 0) & (dz < DW_RANGE_VERT)
            dh = dh * valid; dz = dz * valid
            zero = dz == 0
            dz = np.where(zero, 1.0, dz)
            fdz = -K_D1 * (RP / 4 / dz) ** 2 * np.exp(-0.5 * (dh / (K_D2 * dz + K_D3)) ** 2)
            fdz = np.where(zero, 0.0, fdz)
            f_i[:, 2] += fdz.sum(1)
        u = np.concatenate([f_i, tt[:, 1:4]], 1)
        x = np.concatenate([s[:, 3:6], s[:, 13:16], s[:, 9:13], s[:, 19:22]], 1)
        k1 = _rigid_body(x, u); k2 = _rigid_body(x + dt / 2.0 * k1, u); k3 = _rigid_body(x + dt / 2.0 * k2, u); k4 = _rigid_body(x + dt * k3, u)
        x = x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)  # tools/ode.py:22-31
        x[:, 6:10] /= np.sqrt((x[:, 6:10] ** 2).sum(1, keepdims=True))
        s[:, 3:6], s[:, 13:16], s[:, 9:13], s[:, 19:22] = x[:, 0:3], x[:, 3:6], x[:, 6:10], x[:, 10:13]
        # _update_other_states (:47-61): euler from the NEW quaternion, pdot with the OLD rotation
        e0, e1, e2, e3 = (s[:, 9 + i] for i in range(4))
        s[:, 6] = np.arctan2(2.0 * (e0 * e1 + e2 * e3), e0**2 + e3**2 - e1**2 - e2**2)
        with np.errstate(invalid="ignore", divide="ignore"):
            s[:, 7] = np.arcsin(2.0 * (e0 * e2 - e1 * e3))
            s[:, 8] = np.arctan2(2.0 * (e0 * e3 + e1 * e2), e0**2 + e1**2 - e2**2 - e3**2)
            pdot = np.einsum("nij,nj->ni", R, s[:, 16:19])
            s[:, 23] = np.sqrt((pdot**2).sum(1))
            s[:, 26] = np.arcsin(pdot[:, 2] / s[:, 23])
            s[:, 27] = np.arctan2(pdot[:, 1], pdot[:, 0])
        s[:, 31:35] = delta
        return s

    def forward(self, ts_sim, s, cmd):
        """mul_quadrotors.py:40-50."""
        if self.ctl_t > self.ts_ctl:
            self.delta = self.autopilot(s, cmd, self.all_sim_t)
            self.ctl_t = 0.0
        s = self.dynamics(ts_sim, s, self.delta)
        self.ctl_t += ts_sim
        self.all_sim_t += ts_sim
        return s
