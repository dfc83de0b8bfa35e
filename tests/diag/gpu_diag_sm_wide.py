"""Diagnostic: the SM-wide nominal kernel (NDP_RTI_SM=1) against the 64-thread-CTA kernel (NDP_RTI_SM=0) -- results bit for
bit, and the time of one launch -- on the benchmark workload, a multi-pass batch, a partly filled batch and the stress variant."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.solver import Engine

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(B, scale, stress, steps=3):
    w = wl.independent_problems(B, seed=3, scale=scale)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
    x0, xr, ur = t(w["x0"]), t(w["xr"]), t(w["ur"])
    f = t(np.random.default_rng(4).normal(size=(B, 21, 3)) * (1.0 if stress else 0.1))
    kw = dict(u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0]) if stress else {}
    out = []
    for mode in ("0", "1"):
        os.environ["NDP_RTI_SM"] = mode
        eng = Engine(batch=B, np_=7, precision="f32", **kw)
        eng.reset(xr, ur)
        us = []
        for s in range(steps):
            us.append(eng.update(x0, xr, ur, f).clone())
        torch.cuda.synchronize()
        X, U = eng.get_all("x").clone(), eng.get_all("u").clone()
        st, stats = eng.status().clone(), eng.stats().clone()
        eng.kernel_timing(True)
        ms = []
        for s in range(12):
            eng.reset(xr, ur)
            flush.zero_()
            eng.update(x0, xr, ur, f)
            torch.cuda.synchronize()
            ms.append(eng.last_kernel_ms()[0])
        out.append((us, X, U, st, stats, float(np.median(ms[2:]))))
    a, b = out
    same = all(torch.equal(p, q) for p, q in zip(a[0], b[0])) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    dmax = max(float((p - q).abs().max()) for p, q in zip(a[0], b[0]))
    print("B=%6d scale %.0f stress %d: bit-identical %s (max |du0| %.2e), status!=0 %d / %d, constrained %d; nominal kernel %.2f us (CTA of 64)  %.2f us (SM-wide)"
          % (B, scale, stress, same, dmax, int((a[3] != 0).sum()), int((b[3] != 0).sum()), int((b[4][:, 0] > 1).sum()), a[5] * 1e3, b[5] * 1e3), flush=True)


cases = ((4096, 1.0, 0), (4096, 5.0, 1), (3000, 1.0, 0), (5000, 1.0, 0), (12000, 2.0, 0))
for B, scale, stress in cases[: int(sys.argv[1]) if len(sys.argv) > 1 else len(cases)]:
    run(B, scale, stress)
