"""CPU tests of the host-side mirror of the reference interface (no kernels involved)."""
import numpy as np

from conftest import golden
from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.hv_throttle_est import AlphaFilter, HoverThrottleEstimator
from ndp_nmpc_qd_b200.hv_throttle_est.hover_throttle_estimator import nmpc_u_to_thrust
from ndp_nmpc_qd_b200.params import estimator_params as EP, nmpc_params as CP
from oracle import nmpc_numpy as on


def test_params_match_reference():
    # SURVEY.md B.3
    assert CP.long_list_size == 101 and CP.th_pred == 0.1
    assert CP.xr_list_index == slice(0, 101, 5)
    assert abs(CP.c_max - 27.250000000000004) < 1e-12
    assert EP.k_throttle_init == 50.0 and EP.R == 1.225


def test_hover_throttle_known_answers():
    # constant v_z = 0, hover throttle
    est = HoverThrottleEstimator(0.02)
    ks = [est.update(0.0, 0.27434003169930943)[0] for _ in range(500)]
    assert np.allclose(ks[:5], [50.0806430971306, 50.16480901705952, 50.251776852227835, 50.34085734233125, 50.431401393549066], atol=1e-11)
    assert abs(ks[-1] - 53.0799822023583) < 1e-9


def test_hover_throttle_vs_reference_sequence():
    """Scalar and batched estimators against the reference class driven by the same sequence."""
    g = golden("hv_throttle_golden.npz")
    est = HoverThrottleEstimator(0.02)
    estb = HoverThrottleEstimator(0.02, batch=3)
    for i in range(len(g["vz"])):
        k, x, P = est.update(float(g["vz"][i]), float(g["thr"][i]))
        kb, xb, Pb = estb.update(np.full(3, g["vz"][i]), np.full(3, g["thr"][i]))
        assert abs(k - g["k"][i]) < 1e-9
        assert np.abs(x[:, 0] - g["x"][i]).max() < 1e-9
        assert np.abs(kb - g["k"][i]).max() < 1e-9
    assert x.shape == (2, 1) and P.shape == (2, 2)
    assert np.abs(P - g["P_last"]).max() < 1e-10
    assert np.abs(Pb[1] - g["P_last"]).max() < 1e-10


def test_alpha_filter_and_thrust_map():
    f = AlphaFilter(alpha=0.8, y0=1.0)
    assert abs(f.update(2.0) - 1.2) < 1e-15
    assert nmpc_u_to_thrust(9.81, 50.0) == 9.81 * 1.4844 / 50.0  # nmpc_node.py:281
    assert nmpc_u_to_thrust(9.81, 0.0) == 0.0


def test_reference_horizons_are_dynamically_consistent():
    """The flatness map gives horizons that satisfy the OCP dynamics: integrating (xr_k, ur_k) lands
    near xr_{k+1} (inputs are sampled, so only to O(h^2))."""
    p = on.OcpParams()
    for name in ("eight_high_dyn", "eight_low"):
        xr, ur = wl.reference_horizon([1.0, 4.0], name=name)
        assert np.allclose(np.linalg.norm(xr[..., 6:10], axis=-1), 1.0, atol=1e-12)
        assert np.all(xr[..., 6] > 0)
        for b in range(2):
            for k in range(20):
                xn, _, _ = on.rk4_sens(xr[b, k], ur[b, k], np.zeros(3), p)
                assert np.abs(xn - xr[b, k + 1]).max() < (0.12 if name == "eight_high_dyn" else 2e-3)
    v = wl.figure_eight(np.linspace(0, 11, 500), "eight_high_dyn")[1]
    assert 8.0 < np.linalg.norm(v, axis=-1).max() < 11.0  # envelope of cmd_pc/path_config/eight_high_dyn.yaml


def test_independent_problems_shapes():
    w = wl.independent_problems(16, seed=0, with_neighbour=True)
    assert w["x0"].shape == (16, 10) and w["xr"].shape == (16, 21, 10) and w["ur"].shape == (16, 20, 4)
    assert w["other"].shape == (16, 21, 10)
    d = w["other"][:, 0, 0:2] - w["xr"][:, 0, 0:2]
    assert np.all(np.sum(d * d, 1) < 1.0)
