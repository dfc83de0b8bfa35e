"""Coupled swarm step (BASELINE config 4) on one GPU through SwarmStep: gated all-pairs forces + local NDP-NMPC
solves against the CPU oracles (the multi-GPU exchange is checked by tests/diag/swarm_multi_gpu.py and, for the host
logic, by tests/test_swarm_host.py)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import mlp_numpy
from oracle.c_oracle import make_cfg
from ndp_nmpc_qd_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def test_swarm_step_vs_oracles(built_lib, c_oracle, mlp_weights):
    from ndp_nmpc_qd_b200.swarm import SwarmStep

    n_side, N = 12, 20
    n_all = n_side * n_side
    rng = np.random.default_rng(17)
    # phase-shifted eight_low references on a 0.8 m lattice, altitudes spread over 3 m
    t0 = rng.uniform(0.0, 20.0, n_all)
    xr, ur = wl.reference_horizon(t0, name="eight_low")
    off = np.stack([(np.arange(n_all) % n_side) * 0.8, (np.arange(n_all) // n_side) * 0.8, rng.uniform(0.0, 3.0, n_all)], 1)
    xr = xr.copy()
    xr[:, :, 0:3] += off[:, None, :]
    x0 = xr[:, 0] + 0.02 * rng.normal(size=(n_all, 10))
    x0[:, 6:10] /= np.linalg.norm(x0[:, 6:10], axis=1, keepdims=True)
    sw = SwarmStep(n_all, N=N, mode="local")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    dxr, dur = t(xr), t(ur)
    sw.engine.reset(dxr, dur)
    u0 = sw.step(t(x0), dxr, dur).cpu().numpy().astype(np.float64)
    f = sw.f.cpu().numpy().astype(np.float64)
    torch.cuda.synchronize()
    f_ref = mlp_numpy.swarm_forces(mlp_weights, xr[:, :, 0:6].astype(np.float32), 0, n_all)
    assert np.abs(f_ref).max() > 1.0                      # the lattice does couple
    assert np.abs(f - f_ref).max() < 2e-4
    X, U = xr.copy(), ur.copy()
    r = c_oracle.rti_batch(make_cfg(), x0, xr, ur, f, X, U)
    ok = r["status"] == 0
    st = sw.engine.status().cpu().numpy()
    assert ok.mean() > 0.98 and np.all(st[ok] == 0)
    assert rel_err(u0[ok], r["u0"][ok]) < 1e-4
    assert rel_err(sw.engine.get_all("x").cpu().numpy().astype(np.float64)[ok], X[ok]) < 1e-4


def test_swarm_two_gpus_bit_identical(built_lib, tmp_path):
    """Config 4 on two GPUs (skipped on a one-GPU box): the fused peer-memory exchange, the NCCL all-gather path and the
    single-GPU run must produce the same downwash forces bit for bit, and match the CPU oracle (tests/diag/swarm_multi_gpu.py
    --check asserts p2p == all-gather and |f - oracle| < 1e-4 on every rank)."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tests", "diag", "swarm_multi_gpu.py")
    common = ["--quads", "300", "--steps", "4", "--warmup", "3", "--repeat", "1", "--check", "--dump", str(tmp_path)]
    r1 = subprocess.run([sys.executable, script] + common, capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stdout + r1.stderr
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                         "--master-port", "29641", script] + common, capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    single = np.load(tmp_path / "f_local_w1_r0.npy")
    for mode in ("p2p", "allgather"):
        both = np.concatenate([np.load(tmp_path / f"f_{mode}_w2_r{r}.npy") for r in (0, 1)])
        assert np.array_equal(both, single), mode
