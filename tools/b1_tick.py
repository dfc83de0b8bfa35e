"""Batch-1 fused tick (DownwashNN + controller.update through ndp_pipeline at depth 1): host-side time split."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
from ndp_nmpc_qd_b200.pipeline import HostStepPipeline
from ndp_nmpc_qd_b200.solver import Engine

dev = torch.device("cuda", 0)
nn = DownwashNN(device=dev)
eng = Engine(batch=1, N=20, np_=7, precision="f32", device=dev)
pipe = HostStepPipeline(eng, nn, depth=1)
xr, ur = wl.reference_horizon([1.0])
t32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
eng.reset(t32(xr), t32(ur))
sl = pipe.slots[0]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
fill, sub, wait = [], [], []
for i in range(n):
    xr, ur = wl.reference_horizon([1.0 + 0.02 * i])
    other = xr[0].copy(); other[:, 2] += 0.8
    t0 = time.perf_counter()
    sl.x0[...] = xr[:, 0]; sl.xr[...] = xr; sl.ur[...] = ur
    sl.other[...] = other[None, :, 0:6]; sl.gate_xy[...] = xr[:, 0, 0:2]
    t1 = time.perf_counter()
    pipe.submit(0)
    t2 = time.perf_counter()
    pipe.wait(0)
    t3 = time.perf_counter()
    if i >= 20:
        fill.append(t1 - t0); sub.append(t2 - t1); wait.append(t3 - t2)
print("fill %.1f us  submit %.1f us  wait %.1f us" % (np.median(fill) * 1e6, np.median(sub) * 1e6, np.median(wait) * 1e6))
