import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle.c_oracle import COracle, make_cfg
from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.solver import Engine
co = COracle()
kw = dict(u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0])
for seed, scale, B in ((31, 5.0, 256), (32, 5.0, 2048), (33, 10.0, 2048)):
  for prec in ("f32", "f64"):
    w = wl.independent_problems(B, seed=seed, scale=scale)
    e = Engine(batch=B, np_=4, precision=prec, **kw)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=e.dtype, device="cuda")
    xr, ur = t(w["xr"]), t(w["ur"])
    e.reset(xr, ur); e.set_reference(xr, ur, None)
    u0 = e.solve(t(w["x0"])).cpu().numpy(); torch.cuda.synchronize()
    X, U = w["xr"].copy(), w["ur"].copy()
    r = co.rti_batch(make_cfg(**kw), w["x0"], w["xr"], w["ur"], None, X, U)
    st = e.status().cpu().numpy(); stats = e.stats().cpu().numpy()
    ok = (r["status"] == 0) & (st == 0)
    eu = np.abs(u0 - r["u0"]).max(1) / np.maximum(np.abs(r["u0"]).max(1), 1)
    eU = np.abs(e.get_all("u").cpu().numpy() - U).reshape(B, -1).max(1) / np.maximum(np.abs(U).reshape(B, -1).max(1), 1)
    print(seed, scale, prec, "gpu status", np.bincount(st, minlength=5), "oracle", np.bincount(r["status"], minlength=5), "eu %.2e eU %.2e" % (eu[ok].max(), eU[ok].max()),
          "fact mean %.1f max %d  ipm mean %.1f max %d pol mean %.2f max %d" % (stats[:,0].mean(), stats[:,0].max(), stats[:,1].mean(), stats[:,1].max(), stats[:,2].mean(), stats[:,2].max()))
    for b in np.nonzero(st != 0)[0][:8]:
        print("   bad b", b, "st", st[b], "stats", stats[b], "oracle it", r["n_iter"][b], "nact", r["n_active"][b], "ost", r["status"][b], "eu %.2e" % eu[b])
    bad = np.nonzero(ok & (np.maximum(eu, eU) > 1e-4))[0]
    for b in bad[:5]:
        print("   inaccurate b", b, "stats", stats[b], "nact", r["n_active"][b], "eu %.2e eU %.2e" % (eu[b], eU[b]))
