"""OCP constants (reference: params/nmpc_params.py:5-43; same names so callers keep working)."""
from . import fhnp_params as QD

gravity = QD.gravity
mass = QD.mass

N_node = 20
T_horizon = 2
ts_nmpc = 0.02  # 50 Hz control period
th_pred = T_horizon / N_node

n_states = 10
n_controls = 4

w_max, w_min = 6, -6
c_max, c_min = QD.c_max, 0
v_max, v_min = 20, -20

Qp_xy, Qp_z = 300, 400
Qv_xy, Qv_z = 10, 10
Qq_xy, Qq_z = 10, 100
Rw, Rc = 10, 5

_ratio = th_pred * N_node / ts_nmpc
long_list_size = int(_ratio) + 1
if _ratio - int(_ratio) > 1e-6:
    raise ValueError("please check: th_pred must be an integer multiple of th_nmpc")
xr_list_index = slice(0, long_list_size, int(th_pred / ts_nmpc))
