import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle.c_oracle import COracle, make_cfg
from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.solver import Engine
co = COracle()
for N, prec in ((80, "f32"), (80, "f64"), (40, "f32")):
    B = 64
    w = wl.independent_problems(B, N=N, seed=51, scale=3.0)
    e = Engine(batch=B, N=N, np_=4, precision=prec)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=e.dtype, device="cuda")
    xr, ur = t(w["xr"]), t(w["ur"])
    e.reset(xr, ur); e.set_reference(xr, ur, None)
    u0 = e.solve(t(w["x0"])).cpu().numpy(); torch.cuda.synchronize()
    X, U = w["xr"].copy(), w["ur"].copy()
    r = co.rti_batch(make_cfg(N=N), w["x0"], w["xr"], w["ur"], None, X, U)
    st = e.status().cpu().numpy(); stats = e.stats().cpu().numpy()
    eu = np.abs(u0 - r["u0"]).max(1) / np.maximum(np.abs(r["u0"]).max(1), 1)
    eX = np.abs(e.get_all("x").cpu().numpy() - X).reshape(B, -1).max(1) / np.maximum(np.abs(X).reshape(B, -1).max(1), 1)
    print(N, prec, "max eu %.2e eX %.2e" % (eu.max(), eX.max()))
    for b in np.argsort(-np.maximum(eu, eX))[:6]:
        print("  b", b, "eu %.2e eX %.2e" % (eu[b], eX[b]), "gpu stats", stats[b], "st", st[b], "oracle it", r["n_iter"][b], "nact", r["n_active"][b], "ost", r["status"][b])
