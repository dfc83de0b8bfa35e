"""Diagnostic: two builds of the library (NDP_NMPC_LIB) on the stress set -- dump u0 / iterate / statistics, or compare two dumps bit for bit.
  NDP_NMPC_LIB=a.so python tests/diag/gpu_diag_ab_identical.py dump a.npz;  ... dump b.npz;  python ... cmp a.npz b.npz"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np

if sys.argv[1] == "cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in a.files:
        same = np.array_equal(a[k], b[k], equal_nan=True)
        d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max()
        print("%-12s identical %s  max diff %.3g  (differing entries %d of %d)" % (k, same, d, int((a[k] != b[k]).sum()), a[k].size))
    sys.exit(0)

import torch

from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.solver import Engine

out = {}
for tag, prec, kw in (("f32", "f32", {}), ("f32_ipm", "f32", dict(active_set_first=0)), ("f64", "f64", {})):
    B, N = 2048, 20
    w = wl.independent_problems(B, N=N, seed=5, scale=5.0)
    fd = np.random.default_rng(6).normal(size=(B, N + 1, 3))
    DT = torch.float32 if prec == "f32" else torch.float64
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=DT, device="cuda")
    x0, xr, ur, f = t(w["x0"]), t(w["xr"]), t(w["ur"]), t(fd)
    eng = Engine(batch=B, N=N, np_=7, precision=prec, u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0], **kw)
    eng.reset(xr, ur)
    u0 = torch.empty((B, 4), dtype=DT, device="cuda")
    for s in range(2):  # second step: warm-started
        eng.update(x0, xr, ur, f, u0)
    torch.cuda.synchronize()
    out[tag + "_u0"] = u0.cpu().numpy()
    out[tag + "_X"] = eng.get_all("x").cpu().numpy() if hasattr(eng, "get_all") else np.zeros(1)
    out[tag + "_stats"] = eng.stats().cpu().numpy()
    out[tag + "_status"] = eng.status().cpu().numpy()
    del eng
np.savez(sys.argv[2], **out)
