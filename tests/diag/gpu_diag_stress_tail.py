"""Diagnostic (library built with -DNDP_RTI_PROF): the constrained kernel on the stress variant of config 3 -- when each
problem starts / ends (globaltimer), and what the slowest ones spent their sweeps on."""
import ctypes as C
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ndp_nmpc_qd_b200 import _lib, workloads as wl
from ndp_nmpc_qd_b200.solver import Engine

B, N = 4096, 20
kw = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    kw[k] = int(v)
w = wl.independent_problems(B, N=N, seed=5, scale=5.0)
fd = np.random.default_rng(6).normal(size=(B, N + 1, 3))
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
x0, xr, ur, f = t(w["x0"]), t(w["xr"]), t(w["ur"]), t(fd)
eng = Engine(batch=B, N=N, np_=7, precision="f32", u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0], **kw)
ms = []
for s in range(5):
    eng.reset(xr, ur)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.update(x0, xr, ur, f); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
st = eng.stats().cpu().numpy()
buf = (C.c_ulonglong * (2 * 8192 + 2))()
_lib.load().ndp_debug_rti_cprof(buf)
a = np.array(buf[:], dtype=np.float64)
t0 = a[16384]
start, end = (a[0:2 * B:2] - t0) * 1e-3, (a[1:2 * B:2] - t0) * 1e-3
con = a[0:2 * B:2] > 0
print("settings", kw, "update ms (last)", round(ms[-1], 3), "constrained problems", int(con.sum()))
print("sweeps: mean %.2f max %d;  ipm share %.3f" % (st[:, 0].mean(), st[:, 0].max(), (st[:, 1] > 0).mean()))
e = end[con]
print("problem end times us: p50 %.0f p90 %.0f p99 %.0f p99.9 %.0f max %.0f" % tuple(np.percentile(e, q) for q in (50, 90, 99, 99.9, 100)))
print("problem start times us: p50 %.0f p90 %.0f p99 %.0f max %.0f" % tuple(np.percentile(start[con], q) for q in (50, 90, 99, 100)))
dur = (end - start)[con]
sw = st[con, 0]
print("us per sweep: overall %.1f; problems ending in the last 20%% of the kernel: %.1f" % (dur.sum() / sw.sum(), dur[e > 0.8 * e.max()].sum() / max(1, sw[e > 0.8 * e.max()].sum())))
idx = np.nonzero(con)[0][np.argsort(-e)[:15]]
print("slowest problems (prob, start, end, sweeps, ipm iters, rounds):")
for i in idx:
    print("  %5d %7.0f %7.0f  %3d %3d %3d" % (i, start[i], end[i], st[i, 0], st[i, 1], st[i, 2]))
h = np.bincount(st[con, 0])
print("sweep histogram (count by sweeps):", {int(k): int(v) for k, v in enumerate(h) if v})
