import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.solver import Engine
B, N = 4096, 20
w = wl.independent_problems(B, N=N, seed=5, scale=5.0)
fd = np.random.default_rng(6).normal(size=(B, N + 1, 3))
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
x0, xr, ur, f = t(w["x0"]), t(w["xr"]), t(w["ur"]), t(fd)
eng = Engine(batch=B, N=N, np_=7, precision="f32", u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0])
eng.reset(xr, ur); eng.update(x0, xr, ur, f); torch.cuda.synchronize()
st = eng.stats().cpu().numpy()
ipm = st[:, 1]
print("ipm problems", int((ipm > 0).sum()))
print("ipm iterations histogram", {int(k): int(v) for k, v in enumerate(np.bincount(ipm)) if v and k > 0})
sel = ipm > 0
print("rounds among ipm problems histogram", {int(k): int(v) for k, v in enumerate(np.bincount(st[sel, 2])) if v})
print("fact sweeps among ipm problems", {int(k): int(v) for k, v in enumerate(np.bincount(st[sel, 0])) if v})
