"""Diagnostic: device time of one nominal solve for small batches (the latency build of rti_step_kernel)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.solver import Engine

for B in (1, 128, 1024):
    w = wl.independent_problems(B, N=20, seed=3)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    x0, xr, ur = t(w["x0"]), t(w["xr"]), t(w["ur"])
    eng = Engine(batch=B, N=20, precision="f32")
    eng.reset(xr, ur)
    u0 = torch.empty((B, 4), dtype=torch.float32, device="cuda")
    ms = []
    for s in range(60):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.update(x0, xr, ur, None, u0); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print("B=%d: nominal solve p50 %.1f us (constrained %d)" % (B, np.median(ms[10:]) * 1e3, int((eng.stats().cpu().numpy()[:, 0] > 1).sum())))
