"""GPU tests of the host-buffer C-ABI entry points (ndp_pipeline_*, include/ndp_nmpc.h): one control
step of the whole batch from pinned host memory -- DownwashNN.update + controller.update as the ROS
node calls them per tick (nmpc_node.py:202-209, ndp_nmpc_leader_node.py:60-76)."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import mlp_numpy
from oracle.c_oracle import make_cfg
from ndp_nmpc_qd_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _fill(sl, w):
    sl.x0[...] = w["x0"]; sl.xr[...] = w["xr"]; sl.ur[...] = w["ur"]
    if sl.other is not None:
        sl.other[...] = w["other"][:, :, 0:6]; sl.gate_xy[...] = w["xr"][:, 0, 0:2]


@pytest.mark.parametrize("precision,tol", [("f32", 1e-4), ("f64", 2e-6)])  # f64 solve, but the forces come from the fp32-accurate MLP (~3e-6 N)
def test_pipeline_step_matches_oracle(built_lib, c_oracle, mlp_weights, precision, tol):
    import torch
    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.pipeline import HostStepPipeline
    from ndp_nmpc_qd_b200.solver import Engine

    B = 96
    w = wl.independent_problems(B, seed=11, scale=2.0, with_neighbour=True)
    eng = Engine(batch=B, np_=7, precision=precision)
    nn = DownwashNN()
    pipe = HostStepPipeline(eng, nn, depth=2)
    dt = torch.float32 if precision == "f32" else torch.float64
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
    eng.reset(t(w["xr"]), t(w["ur"]))
    torch.cuda.synchronize()
    _fill(pipe.slots[0], w)
    out = pipe.step(0)
    ndt = np.float32 if precision == "f32" else np.float64
    f_ref = mlp_numpy.gated_pairs(mlp_weights, w["xr"].astype(ndt).astype(np.float64), w["other"].astype(ndt).astype(np.float64),
                                  w["xr"][:, 0, 0:2].astype(ndt).astype(np.float64))
    X, U = w["xr"].copy(), w["ur"].copy()
    r = c_oracle.rti_batch(make_cfg(), w["x0"], w["xr"], w["ur"], f_ref, X, U)
    ok = r["status"] == 0
    assert ok.sum() >= B - 2
    assert (out.status[ok] == 0).all()
    assert rel_err(out.u0[ok], r["u0"][ok]) < tol
    assert pipe.h2d_bytes_per_step == B * (10 + 210 + 80 + 126 + 2) * np.dtype(ndt).itemsize
    assert pipe.d2h_bytes_per_step == B * 4 * np.dtype(ndt).itemsize + B * 4


def test_pipelined_steps_equal_sequential_device_path(built_lib):
    """depth-3 overlapped submission of 7 consecutive control steps == the same steps through the
    device-pointer API one at a time (bit-identical u0: same kernels, same order)."""
    import torch
    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.pipeline import HostStepPipeline
    from ndp_nmpc_qd_b200.solver import Engine

    B, steps, depth = 300, 7, 3
    base = wl.independent_problems(B, seed=5, with_neighbour=True)
    rng = np.random.default_rng(0)
    x0s = [(base["x0"] + 0.02 * rng.normal(size=base["x0"].shape)).astype(np.float32) for _ in range(steps)]
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    nn = DownwashNN()
    # sequential reference through the device API
    eng = Engine(batch=B, np_=7)
    xr, ur, other, gate = t(base["xr"]), t(base["ur"]), t(base["other"]), t(base["xr"][:, 0, 0:2])
    eng.reset(xr, ur)
    ref = []
    for s in range(steps):
        f = nn.forward_pairs(xr, other, gate)
        ref.append(eng.update(t(x0s[s]), xr, ur, f).cpu().numpy())
    # pipelined
    eng2 = Engine(batch=B, np_=7)
    eng2.reset(xr, ur)
    torch.cuda.synchronize()
    pipe = HostStepPipeline(eng2, nn, depth=depth)
    got = [None] * steps
    for s in range(steps):
        if s >= depth:
            got[s - depth] = pipe.wait(s % depth).u0.copy()
        sl = pipe.slots[s % depth]
        _fill(sl, base)
        sl.x0[...] = x0s[s]
        pipe.submit(s % depth)
    for s in range(steps - depth, steps):
        got[s] = pipe.wait(s % depth).u0.copy()
    for s in range(steps):
        assert np.array_equal(got[s], ref[s]), s
    # depth 1 = the latency path: the step is captured once into a CUDA graph and replayed (same kernels, same order)
    eng3 = Engine(batch=B, np_=7)
    eng3.reset(xr, ur)
    torch.cuda.synchronize()
    pipe1 = HostStepPipeline(eng3, nn, depth=1)
    for s in range(steps):
        sl = pipe1.slots[0]
        _fill(sl, base)
        sl.x0[...] = x0s[s]
        assert np.array_equal(pipe1.step(0).u0, ref[s]), s
        assert np.all(sl.status == 0)


def test_pipeline_without_downwash(built_lib, c_oracle):
    import torch
    from ndp_nmpc_qd_b200.pipeline import HostStepPipeline
    from ndp_nmpc_qd_b200.solver import Engine

    B = 33
    w = wl.independent_problems(B, seed=3)
    eng = Engine(batch=B, np_=4)
    pipe = HostStepPipeline(eng, None, depth=1)
    assert pipe.slots[0].other is None
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    eng.reset(t(w["xr"]), t(w["ur"]))
    torch.cuda.synchronize()
    _fill(pipe.slots[0], w)
    out = pipe.step(0)
    X, U = w["xr"].copy(), w["ur"].copy()
    r = c_oracle.rti_batch(make_cfg(), w["x0"], w["xr"], w["ur"], None, X, U)
    assert rel_err(out.u0, r["u0"]) < 1e-4 and (out.status == 0).all()


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_device_long_list_reproduces_the_reference_publisher(built_lib, precision):
    """ndp_longlist_push against NMPCRefPublisher itself (pt_pub/pt_publisher.py:57-103; tests/golden/wire_golden.npz holds
    its 101-point list after reset(), the point each of 130 get_nmpc_pts() calls appended and the (xr, ur) each returned):
    the device rings must hand back exactly those horizons (they only move data)."""
    import torch
    from conftest import golden
    from ndp_nmpc_qd_b200.pipeline import LongList
    from ndp_nmpc_qd_b200.solver import Engine

    g = golden("wire_golden.npz")
    B = 3  # the same list three times, with a per-problem shift so that a mixed-up problem index shows
    shift = np.arange(B, dtype=np.float64)[:, None, None]
    eng = Engine(batch=B, np_=7, precision=precision)
    ll = LongList(eng, with_other=True)
    ndt = np.float32 if precision == "f32" else np.float64
    x_init = (g["longlist_x_init"][None] + shift).astype(ndt)
    u_init = (g["longlist_u_init"][None] + shift).astype(ndt)
    ll.reset(x_init, u_init, x_init[:, :, 0:6].copy())
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=eng.dtype, device="cuda")
    for j in range(g["longlist_new_x"].shape[0]):
        nx = (g["longlist_new_x"][j][None] + shift[:, 0]).astype(ndt)
        nu = (g["longlist_new_u"][j][None] + shift[:, 0]).astype(ndt)
        xr, ur, other = ll.push(t(nx), t(nu), t(nx[:, 0:6]))
        ref_x = (g["longlist_xr"][j + 1][None] + shift).astype(ndt)
        ref_u = (g["longlist_ur"][j + 1][None] + shift).astype(ndt)
        assert np.array_equal(xr.cpu().numpy(), ref_x), j
        assert np.array_equal(ur.cpu().numpy(), ref_u), j
        assert np.array_equal(other.cpu().numpy(), ref_x[:, :, 0:6]), j


def test_long_list_pipeline_equals_full_horizon_pipeline(built_lib):
    """The step pipeline fed ONE new point per list and step (long-list mode, 128 B per problem) against the pipeline fed
    the whole horizons (1 712 B per problem): same kernels on the same horizons -> bit-identical u0 over 9 steps."""
    import torch
    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.pipeline import HostStepPipeline, LongList
    from ndp_nmpc_qd_b200.solver import Engine

    B, steps, depth = 200, 9, 3
    rng = np.random.default_rng(5)
    t0 = rng.uniform(0.0, 9.0, B)
    # ego lists at 50 Hz: point i of tick j = reference at t0 + 0.02 (i + j); the neighbour flies 0.8 m above
    tt = t0[:, None] + 0.02 * np.arange(101 + steps)[None, :]
    from ndp_nmpc_qd_b200.workloads import diff_flatness, figure_eight
    xl, ul = diff_flatness(*figure_eight(tt, "eight_high_dyn"))
    ol = xl[:, :, 0:6].copy(); ol[:, :, 2] += 0.8
    eng_a, eng_b = Engine(batch=B, np_=7), Engine(batch=B, np_=7)
    nn = DownwashNN()
    ll = LongList(eng_b, with_other=True)
    ll.reset(xl[:, :101], ul[:, :101], ol[:, :101])
    pa, pb = HostStepPipeline(eng_a, nn, depth=depth), HostStepPipeline(eng_b, nn, depth=depth, longlist=ll)
    assert pb.h2d_bytes_per_step == B * (10 + 10 + 4 + 6 + 2) * 4 and pa.h2d_bytes_per_step == B * (10 + 210 + 80 + 126 + 2) * 4
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    for e in (eng_a, eng_b):
        e.reset(t(xl[:, 0:101:5]), t(ul[:, 0:100:5]))
    torch.cuda.synchronize()
    ua, ub = [], []
    for s in range(steps):
        j = s + 1  # tick j: list = points j .. j + 100
        x0 = xl[:, j] + 0.01 * rng.normal(size=(B, 10)) * np.array([1, 1, 1, 1, 1, 1, 0, 0, 0, 0])
        sa, sb = pa.slots[s % depth], pb.slots[s % depth]
        if s >= depth:
            ua.append(pa.wait(s % depth).u0.copy()); ub.append(pb.wait(s % depth).u0.copy())
        sa.x0[...] = x0; sa.xr[...] = xl[:, j:j + 101:5]; sa.ur[...] = ul[:, j:j + 100:5]
        sa.other[...] = ol[:, j:j + 101:5]; sa.gate_xy[...] = xl[:, j, 0:2]
        sb.x0[...] = x0; sb.xr[:, 0] = xl[:, j + 100]; sb.ur[:, 0] = ul[:, j + 100]
        sb.other[:, 0] = ol[:, j + 100]; sb.gate_xy[...] = xl[:, j, 0:2]
        pa.submit(s % depth); pb.submit(s % depth)
    for s in range(steps - depth, steps):
        ua.append(pa.wait(s % depth).u0.copy()); ub.append(pb.wait(s % depth).u0.copy())
        assert (pb.slots[s % depth].status == 0).all()
    assert len(ua) == steps and all(np.array_equal(a, b) for a, b in zip(ua, ub))
    assert np.abs(np.stack(ua)).max() > 1.0
