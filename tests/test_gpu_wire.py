"""GPU tests of the batched glue kernels around the controller against
tests/golden/wire_golden.npz / hv_throttle_golden.npz -- outputs of the reference's own node methods and estimator class
(tests/golden/make_wire_golden.py, make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def _t(a, dt=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_odometry_to_x0_kernel(built_lib, dt):
    """dop_qd_node.py:131-148 + pt_publisher.py:106-122 via ndp_plant_nmpc_x0."""
    from ndp_nmpc_qd_b200.dop_sim import MulQuadrotors

    g = golden("wire_golden.npz")
    n = g["odom_state"].shape[0]
    plant = MulQuadrotors(n, 0.01, 0.01, torch.float64, False, True, False)
    x0 = plant.nmpc_x0(_t(g["odom_state"])[:, :, None].contiguous(), torch.empty((n, 10), dtype=dt, device="cuda"))
    ref = g["odom_x0"].astype(np.float32 if dt == torch.float32 else np.float64)
    assert np.array_equal(x0.cpu().numpy(), ref)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_attitude_target_kernels(built_lib, dt):
    """nmpc_node.py:273-283 + dop_qd_node.py:162-166 via ndp_plant_cmd_from_u0 (one k_throttle) and
    ndp_plant_cmd_from_u0_dev (one per quadrotor, incl. the k_throttle == 0 branch)."""
    from ndp_nmpc_qd_b200.dop_sim import MulQuadrotors

    g = golden("wire_golden.npz")
    n = g["att_u0"].shape[0]
    plant = MulQuadrotors(n, 0.01, 0.01, torch.float64, False, True, False)
    u0 = _t(g["att_u0"], dt)
    u0_host = u0.cpu().numpy().astype(np.float64)
    cmd = plant.cmd_from_u0_dev(u0, torch.empty((n, 4, 1), dtype=torch.float64, device="cuda"), 1.4844, _t(g["att_k_throttle"]))
    ref = g["att_cmd"].copy()
    if dt == torch.float32:  # the kernel sees u0 rounded to fp32, then works in fp64 like the node
        k = g["att_k_throttle"]
        ref = np.concatenate([u0_host[:, :3], np.where(k != 0, u0_host[:, 3] * 1.4844 / np.where(k != 0, k, 1.0), 0.0)[:, None]], 1)
    assert np.allclose(cmd.cpu().numpy()[:, :, 0], ref, rtol=1e-15, atol=0)
    assert cmd[2, 3, 0].item() == 0.0
    cmd1 = plant.cmd_from_u0(u0, torch.empty((n, 4, 1), dtype=torch.float64, device="cuda"), 1.4844, 50.0)
    assert np.allclose(cmd1.cpu().numpy()[:, 3, 0], u0_host[:, 3] * 1.4844 / 50.0, rtol=1e-15, atol=0)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_predxu_pack_unpack(built_lib, dt):
    """PredXU payload (msg/PredXU.msg; do_pub_ref nmpc_node.py:116-133) and the follower's consumption of it with the
    formation offset (nmpc_follower_node.py:57-74)."""
    from ndp_nmpc_qd_b200 import wire

    g = golden("wire_golden.npz")
    B = g["predxu_xr"].shape[0]
    assert wire.predxu_len(20) == 290 == g["predxu_msg"].shape[1]
    xr, ur = _t(g["predxu_xr"], dt), _t(g["predxu_ur"], dt)
    msg = wire.predxu_pack(xr, ur)
    assert msg.dtype == torch.float64
    ref_msg = g["predxu_msg"] if dt == torch.float64 else np.concatenate(
        [g["predxu_xr"].astype(np.float32).reshape(B, -1), g["predxu_ur"].astype(np.float32).reshape(B, -1)], 1).astype(np.float64)
    assert np.array_equal(msg.cpu().numpy(), ref_msg)
    fx, fu = wire.predxu_unpack(_t(g["predxu_msg"]), 20, dtype=dt, offset=_t(g["predxu_offset"]))
    np_dt = np.float32 if dt == torch.float32 else np.float64
    assert np.array_equal(fx.cpu().numpy(), g["predxu_follower_xr"].astype(np_dt))
    assert np.array_equal(fu.cpu().numpy(), g["predxu_follower_ur"].astype(np_dt))
    lx, lu = wire.predxu_unpack(_t(g["predxu_msg"]), 20, dtype=dt)  # leader side: no offset (ndp_nmpc_leader_node.py:69-71)
    assert np.array_equal(lx.cpu().numpy(), g["predxu_xr"].astype(np_dt)) and np.array_equal(lu.cpu().numpy(), g["predxu_ur"].astype(np_dt))
    # empty batch
    e = wire.predxu_pack(xr[:0].contiguous(), ur[:0].contiguous())
    assert e.shape == (0, 290)


def test_device_hover_throttle_estimator(built_lib):
    """Batched device filter vs the reference class's recorded run (hv_throttle_golden.npz) and vs the host mirror on
    random per-quadrotor sequences, incl. samples outside the 0.1 < throttle < 1 window."""
    from ndp_nmpc_qd_b200.hv_throttle_est import HoverThrottleEstimator
    from ndp_nmpc_qd_b200.wire import BatchedHoverThrottleEstimator

    g = golden("hv_throttle_golden.npz")
    vz, thr, k_ref = g["vz"], g["thr"], g["k"]
    n = 7
    est = BatchedHoverThrottleEstimator(n)
    host = HoverThrottleEstimator(0.02, batch=n)
    rng = np.random.default_rng(0)
    buf = torch.zeros((n, 35, 1), dtype=torch.float64, device="cuda")   # strided views like the closed loop uses
    cmd = torch.zeros((n, 4, 1), dtype=torch.float64, device="cuda")
    for i in range(len(vz)):
        v = np.concatenate([[vz[i]], vz[i] + 0.05 * rng.normal(size=n - 1)])
        th = np.concatenate([[thr[i]], rng.uniform(0.0, 1.1, size=n - 1)])
        buf[:, 15, 0] = torch.as_tensor(v, device="cuda")
        cmd[:, 3, 0] = torch.as_tensor(th, device="cuda")
        k_dev = est.update(buf[:, 15, 0], cmd[:, 3, 0]).cpu().numpy()
        k_host = host.update(v, th)[0]
        assert abs(k_dev[0] - k_ref[i]) < 1e-10 * max(1.0, abs(k_ref[i])), i
        assert np.allclose(k_dev, k_host, rtol=1e-12, atol=0), i
    est.reset()
    assert np.all(est.k_throttle.cpu().numpy() == 50.0)


def test_closed_loop_with_device_estimator(built_lib):
    """ClosedLoop(k_throttle=None): the hover phase runs the estimator on the device (nmpc_node.py:251-253) from the
    reference's initial guess 50 towards m g / throttle_hover = 53.08, tracking then freezes it."""
    from ndp_nmpc_qd_b200 import traj_gen
    from ndp_nmpc_qd_b200.closed_loop import K_THROTTLE, ClosedLoop

    tr = traj_gen.plan_named("eight_low")
    B = 32
    cl = ClosedLoop([tr], np.zeros(B, np.int32), np.zeros(B), k_throttle=None)
    assert np.all(cl.estimator.k_throttle.cpu().numpy() == 50.0)
    cl.hover(400)
    k = cl.estimator.k_throttle.cpu().numpy()
    assert np.all(np.abs(k - K_THROTTLE) < 0.05), k[:4]
    for _ in range(50):
        cl.step()
    torch.cuda.synchronize()
    assert np.array_equal(cl.estimator.k_throttle.cpu().numpy(), k)  # frozen while tracking (nmpc_node.py:146)
    assert (cl.engine.status().cpu().numpy() == 0).all() and float(cl.position_error().max()) < 0.05


def test_handles_run_on_their_own_device(built_lib, c_oracle):
    """an Engine / DownwashNN built for cuda:1 must work while cuda:0 is the thread's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from ndp_nmpc_qd_b200 import workloads as wl
    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.solver import Engine
    from oracle.c_oracle import make_cfg
    from conftest import rel_err

    torch.cuda.set_device(0)
    B = 64
    w = wl.independent_problems(B, seed=3, with_neighbour=True)
    eng, nn = Engine(batch=B, np_=7, device="cuda:1"), DownwashNN(device="cuda:1")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda:1")
    xr, ur = t(w["xr"]), t(w["ur"])
    assert torch.cuda.current_device() == 0
    f = nn.forward_pairs(xr, t(w["other"]), t(w["xr"][:, 0, 0:2]))
    eng.reset(xr, ur)
    u0 = eng.update(t(w["x0"]), xr, ur, f)
    torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0 and u0.device.index == 1
    X, U = w["xr"].copy(), w["ur"].copy()
    r = c_oracle.rti_batch(make_cfg(), w["x0"], w["xr"], w["ur"], f.cpu().numpy().astype(np.float64), X, U)
    assert rel_err(u0.cpu().numpy(), r["u0"]) < 1e-4
