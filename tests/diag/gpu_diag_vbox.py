"""Diagnostic (not a test): velocity-box problems, per-route accuracy and sweep statistics; lone-problem round latency."""
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
vm = np.array(wl.V_BOX)
kw = dict(v_min=list(-vm), v_max=list(vm))
B = 512
w = wl.velocity_box_problems(B, seed=12)
X, U = w["xr"].copy(), w["ur"].copy()
r = co.rti_batch(make_cfg(**kw), w["x0"], w["xr"], w["ur"], None, X, U)
for prec in ("f32", "f64"):
    for as_first in (16, 6, 0):
        for pol in (12, 30):
            e = Engine(batch=B, np_=4, precision=prec, active_set_first=as_first, polish_max=pol, **kw)
            t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=e.dtype, device="cuda")
            xr, ur = t(w["xr"]), t(w["ur"])
            e.reset(xr, ur); e.set_reference(xr, ur, None)
            u0 = e.solve(t(w["x0"])).cpu().numpy().astype(np.float64); torch.cuda.synchronize()
            st, stats = e.status().cpu().numpy(), e.stats().cpu().numpy()
            eu = (np.abs(u0 - r["u0"]) / np.maximum(np.abs(r["u0"]), 1)).max(1)
            eU = (np.abs(e.get_all("u").cpu().numpy() - U) / np.maximum(np.abs(U), 1)).reshape(B, -1).max(1)
            print(f"{prec} as_first={as_first} polish={pol}: status {np.bincount(st, minlength=5)} eu max {eu.max():.2e} eU max {eU.max():.2e} "
                  f"sweeps mean {stats[:,0].mean():.1f} max {stats[:,0].max()} ipm share {(stats[:,1]>0).mean():.2f} rounds mean {stats[:,2].mean():.1f} max {stats[:,2].max()}")
            for b in np.argsort(-eU)[:3]:
                print(f"    b={b} eu {eu[b]:.2e} eU {eU[b]:.2e} stats {stats[b]} oracle nact {r['n_active'][b]}")
# lone-problem latency of the constrained path
for kwl, name, gen in ((dict(u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0]), "tight-u", lambda: wl.independent_problems(64, seed=31, scale=5.0)),
                       (kw, "v-box", lambda: wl.velocity_box_problems(64, seed=5))):
    ws_ = gen()
    for b in range(6):
        e = Engine(batch=1, np_=4, precision="f32", **kwl)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
        xr, ur, x0 = t(ws_["xr"][b:b+1]), t(ws_["ur"][b:b+1]), t(ws_["x0"][b:b+1])
        ms = []
        for rep in range(6):
            e.reset(xr, ur); e.set_reference(xr, ur, None)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); e.solve(x0); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) * 1e3)
        s = e.stats().cpu().numpy()[0]
        print(f"lone {name} b={b}: {np.median(ms):.1f} us, sweeps {s[0]} ipm {s[1]} rounds {s[2]} nact {s[3]} -> {(np.median(ms)-55)/max(1,s[0]-1):.1f} us per extra sweep")
