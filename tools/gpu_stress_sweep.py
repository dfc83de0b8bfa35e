"""Stress variant of config 3 (bench.py: stress_variant) over the number of active-set rounds tried before the IPM."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ndp_nmpc_qd_b200 import workloads as wl  # noqa: E402
from ndp_nmpc_qd_b200.solver import Engine  # noqa: E402

B, N = 4096, 20
dev = torch.device("cuda", 0)
w = wl.independent_problems(B, N=N, seed=5, scale=5.0)
fd = np.random.default_rng(6).normal(size=(B, N + 1, 3))
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
x0, xr, ur, f = t(w["x0"]), t(w["xr"]), t(w["ur"]), t(fd)
u0 = torch.empty((B, 4), dtype=torch.float32, device=dev)
ref = None
for as_first in [int(a) for a in (sys.argv[1:] or ["0", "6", "10", "16", "24", "40"])]:
    eng = Engine(batch=B, N=N, np_=7, precision="f32", device=dev, u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0],
                 active_set_first=as_first)
    ms = []
    for s in range(8):
        eng.reset(xr, ur)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.update(x0, xr, ur, f, u0); e1.record()
        torch.cuda.synchronize()
        if s >= 2:
            ms.append(e0.elapsed_time(e1))
    st = eng.stats().cpu().numpy()
    u = u0.cpu().numpy().copy()
    if ref is None:
        ref = u
    pol_hist = np.bincount(st[st[:, 1] == 0, 2], minlength=8)[:30].tolist()
    print(json.dumps(dict(as_first=as_first, kernel_ms=float(np.mean(ms)), sweeps_mean=float(st[:, 0].mean()), sweeps_max=int(st[:, 0].max()),
                          ipm_share=float((st[:, 1] > 0).mean()), rounds_hist_no_ipm=pol_hist, status_nonzero=int((eng.status().cpu().numpy() != 0).sum()),
                          u0_max_diff_vs_first=float(np.abs(u - ref).max()))))
    del eng
