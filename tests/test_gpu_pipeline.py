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
