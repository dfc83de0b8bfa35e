"""Diagnostic (not a test): warm vs cold active-set route in a saturating closed loop (the setting of
test_warm_active_set_closed_loop), per-step worst problem."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from oracle.c_oracle import COracle, make_cfg
from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.solver import Engine

co = COracle()
B, steps = 128, 6
kw = dict(u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0])
w = wl.independent_problems(B, seed=61, scale=5.0)
rng = np.random.default_rng(8)
x0_seq = [w["x0"] + 0.03 * s * rng.normal(size=w["x0"].shape) for s in range(steps)]
for prec in ("f32", "f64"):
    ew, ec = Engine(batch=B, np_=4, precision=prec, active_set_warm=1, **kw), Engine(batch=B, np_=4, precision=prec, active_set_warm=0, **kw)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=ew.dtype, device="cuda")
    xr, ur = t(w["xr"]), t(w["ur"])
    X, U = w["xr"].copy(), w["ur"].copy()
    for e in (ew, ec):
        e.reset(xr, ur); e.set_reference(xr, ur, None)
    for s in range(steps):
        r = co.rti_batch(make_cfg(**kw), x0_seq[s], w["xr"], w["ur"], None, X, U)
        for name, e in (("warm", ew), ("cold", ec)):
            e.solve(t(x0_seq[s])); torch.cuda.synchronize()
            st, stats = e.status().cpu().numpy(), e.stats().cpu().numpy()
            gU = e.get_all("u").cpu().numpy().astype(np.float64)
            err = (np.abs(gU - U) / np.maximum(np.abs(U), 1)).reshape(B, -1).max(1)
            b = int(np.argmax(err))
            print(f"{prec} step {s} {name}: max err {err.max():.2e} at b={b} stats {stats[b]} status {st[b]} oracle it {r['n_iter'][b]} nact {r['n_active'][b]}; "
                  f"sweeps mean {stats[:,0].mean():.2f} ipm share {(stats[:,1]>0).mean():.3f}")
