"""CPU tests of the host-side node logic around the controller, against
tests/golden/wire_golden.npz -- outputs of the reference's OWN node methods (tests/golden/make_wire_golden.py)."""
import numpy as np

from conftest import golden
from ndp_nmpc_qd_b200 import formation as fm
from ndp_nmpc_qd_b200.hv_throttle_est import Differentiator, HoverThrottleEstimator


def test_odom_and_attitude_target_mappings():
    g = golden("wire_golden.npz")
    x0 = np.stack([fm.odom_to_x0(s) for s in g["odom_state"]])
    assert np.array_equal(x0, g["odom_x0"])
    cmd = np.stack([fm.u0_to_cmd(u, k) for u, k in zip(g["att_u0"], g["att_k_throttle"])])
    assert np.array_equal(cmd, g["att_cmd"])
    assert cmd[2, 3] == 0.0  # k_throttle == 0 -> thrust 0 (nmpc_node.py:281)


def test_follower_reference_from_predxu():
    g = golden("wire_golden.npz")
    for b in range(g["predxu_xr"].shape[0]):
        xr, ur = fm.follower_reference(g["predxu_xr"][b], g["predxu_ur"][b], g["predxu_offset"][b])
        assert np.array_equal(xr, g["predxu_follower_xr"][b]) and np.array_equal(ur, g["predxu_follower_ur"][b])


def test_formation_reference_switch_and_filter():
    g = golden("wire_golden.npz")
    fx, fs = fm.FormationOffsetFilter(), fm.FormationOffsetFilter()
    for i, xl in enumerate(g["form_leader_x"]):
        xf, sb = fm.leader_formation_refs(float(xl))
        assert np.array_equal(xf, g["form_xf_raw"][i]) and np.array_equal(sb, g["form_sb_raw"][i])
        assert np.allclose(fx.update(xf), g["form_xf_filtered"][i], rtol=0, atol=1e-15)
        assert np.allclose(fs.update(sb), g["form_sb_filtered"][i], rtol=0, atol=1e-15)
    assert len(np.unique(g["form_xf_raw"], axis=0)) == 2  # both branches of the switch are exercised


def test_leader_gate():
    g = golden("wire_golden.npz")
    got = [fm.gate_open(o, e) for o, e in zip(g["gate_other_xy"], g["gate_ego_xy"])]
    assert got == list(g["gate_mlp_called"]) and any(got) and not all(got)
    f = fm.leader_disturb_force(lambda other, ego: np.ones((21, 3)), np.zeros((21, 10)) + 5.0, (0.0, 0.0), np.zeros((21, 10)))
    assert f.shape == (21, 3) and not f.any()  # gate closed -> zeros [21, 3]


def test_differentiator_keeps_a_private_copy():
    """a batched caller that reuses one vz buffer in place must not make x - x_delay_1 vanish."""
    d_ref, d_buf = Differentiator(0.02, batch=3), Differentiator(0.02, batch=3)
    buf = np.zeros(3)
    for k in range(5):
        v = np.array([0.1, -0.2, 0.3]) * (k + 1)
        buf[:] = v
        assert np.array_equal(d_ref.update(v.copy()), d_buf.update(buf))
    est = HoverThrottleEstimator(0.02, batch=3)
    for k in range(5):
        buf[:] = 0.01 * k
        est.update(buf, np.full(3, 0.27))
    assert np.all(est.x[:, 1] != 50.0)
