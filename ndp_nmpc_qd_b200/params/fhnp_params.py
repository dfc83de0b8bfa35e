"""Airframe constants (reference: params/fhnp_params.py:9-45; the hot path reads mass, gravity and c_max, the rest
is kept so that `from params import fhnp_params as AP` in the reference's pt_pub resolves every name it may use)."""
import math

mass = 1.4844  # kg
l_frame = 0.1372  # m
alpha_frame = 45.0 * math.pi / 180.0  # rad
gravity = 9.81  # m/s^2
Jx, Jy, Jz, Jxz = 0.0094, 0.0134, 0.0145, 0.0  # kg m^2
c_max = gravity / 0.36  # max collective acceleration, m/s^2
o_max = 24000 / 1000  # kRPM
o_min = 2600 / 1000  # kRPM
c_q = 3.7611e-10  # Nm/RPM^2
c_t = 2.8158e-08  # N/RPM^2
t_w_r = 4.3100  # thrust-to-weight ratio
# inertia terms of the rigid-body equations
gamma = Jx * Jz - (Jxz**2)
gamma1 = (Jxz * (Jx - Jy + Jz)) / gamma
gamma2 = (Jz * (Jz - Jy) + (Jxz**2)) / gamma
gamma3 = Jz / gamma
gamma4 = Jxz / gamma
gamma5 = (Jz - Jx) / Jy
gamma6 = Jxz / Jy
gamma7 = ((Jx - Jy) * Jx + (Jxz**2)) / gamma
gamma8 = Jx / gamma
