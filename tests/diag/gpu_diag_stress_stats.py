"""Diagnostic: per-problem solver statistics of the stress set (factorisations, IPM iterations, rounds, active bounds) -> npz."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.solver import Engine

B, N = 4096, 20
w = wl.independent_problems(B, N=N, seed=5, scale=5.0)
fd = np.random.default_rng(6).normal(size=(B, N + 1, 3))
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
x0, xr, ur, f = t(w["x0"]), t(w["xr"]), t(w["ur"]), t(fd)
eng = Engine(batch=B, N=N, np_=7, precision="f32", u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0])
eng.reset(xr, ur)
eng.update(x0, xr, ur, f)
torch.cuda.synchronize()
U = eng.get_all("u").cpu().numpy().reshape(B, N, 4)
np.savez(sys.argv[1], stats=eng.stats().cpu().numpy(), U=U)
