"""Diagnostic: one hard problem of the stress set alone on the GPU -- time per sweep of the IPM-first and rounds-first routes."""
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
PREC = sys.argv[1] if len(sys.argv) > 1 else "f32"
DT = torch.float32 if PREC == "f32" else torch.float64
for p in (596, 3722, 1752, 2726):
    sl = slice(p, p + 1)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a[sl]), dtype=DT, device="cuda")
    x0, xr, ur, f = t(w["x0"]), t(w["xr"]), t(w["ur"]), t(fd)
    for kw in (dict(active_set_first=0), dict(active_set_first=8), dict()):
        eng = Engine(batch=1, N=N, np_=7, precision=PREC, u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0], **kw)
        ms = []
        for s in range(6):
            eng.reset(xr, ur)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.update(x0, xr, ur, f); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        st = eng.stats().cpu().numpy()[0]
        print("problem %d %s: %.0f us, sweeps %d, ipm iterations %d, rounds %d -> %.1f us per sweep" % (p, kw, np.median(ms) * 1e3, st[0], st[1], st[2], np.median(ms) * 1e3 / st[0]))

# interior-point phase times of the LAST problem's runs (library built with -DNDP_RTI_PROF: clock64 sums since the
# library was loaded, so the numbers are relative shares)
import ctypes as C
from ndp_nmpc_qd_b200 import _lib
lib = _lib.load()
if hasattr(lib, "ndp_debug_rti_cprof"):
    buf = (C.c_ulonglong * (2 * 8192 + 2))()
    lib.ndp_debug_rti_cprof(buf)
    v = np.array(buf[100:109], dtype=np.float64)
    names = ["rounds->loop gap", "barrier terms loop", "predictor backward", "predictor forward", "predictor step loops", "delta backward",
             "delta forward", "step length + update loops", "rounds from the IPM estimate"]
    for n, x in zip(names, v):
        print("%-32s %6.1f %%" % (n, 100 * x / v.sum()))
    r = np.array(buf[110:113], dtype=np.float64)
    for n, x in zip(["round: backward sweep", "round: forward sweep", "round: multipliers + set update"], r):
        print("%-32s %6.1f %%   %8.0f cycles per round" % (n, 100 * x / r.sum(), x / max(buf[120], 1)))
    print("rounds %d, backward stages per round %.1f" % (buf[120], buf[121] / max(buf[120], 1)))
