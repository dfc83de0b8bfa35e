"""ndp_nmpc_qd_b200 -- B200-native batched NMPC engine behind the controller surface of
Li-Jinjie/ndp_nmpc_qd (see DESIGN.md / INTEGRATION.md).

Sub-packages keep the reference's import names (nmpc_ctl, ndp_nmpc_ctl, dnwash_nn_est,
hv_throttle_est, params; nmpc_node.py:29-32).
"""
__version__ = "0.1.0"
