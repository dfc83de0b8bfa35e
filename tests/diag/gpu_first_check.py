"""Diagnostic run for a GPU box: prints parity numbers without asserting (so one call tells a lot)."""
import sys, os, time, traceback
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import rel_err
from oracle.c_oracle import COracle, make_cfg
from ndp_nmpc_qd_b200 import workloads as wl
from ndp_nmpc_qd_b200.solver import Engine

co = COracle()
print(torch.cuda.get_device_name(0))
for prec in ("f32", "f64"):
    for scale, B in ((1.0, 512), (5.0, 512), (15.0, 512)):
        try:
            w = wl.independent_problems(B, seed=21, scale=scale)
            fd = np.random.default_rng(4).normal(size=(B, 21, 3))
            e = Engine(batch=B, np_=7, precision=prec)
            t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=e.dtype, device="cuda")
            xr, ur = t(w["xr"]), t(w["ur"])
            e.reset(xr, ur); e.set_reference(xr, ur, t(fd))
            u0 = e.solve(t(w["x0"])); torch.cuda.synchronize()
            X, U = w["xr"].copy(), w["ur"].copy()
            r = co.rti_batch(make_cfg(), w["x0"], w["xr"], w["ur"], fd, X, U)
            st = e.status().cpu().numpy(); stats = e.stats().cpu().numpy()
            ok = (r["status"] == 0) & (st == 0)
            eu = np.abs(u0.cpu().numpy() - r["u0"]).max(1) / np.maximum(np.abs(r["u0"]).max(1), 1)
            eX = np.abs(e.get_all("x").cpu().numpy() - X).reshape(B, -1).max(1) / np.maximum(np.abs(X).reshape(B, -1).max(1), 1)
            print(prec, scale, "status gpu", np.bincount(st, minlength=5), "oracle", np.bincount(r["status"], minlength=5),
                  "u0 rel max %.3e (ok) %.3e (all)" % (eu[ok].max() if ok.any() else -1, eu.max()), "X rel max %.3e" % (eX[ok].max() if ok.any() else -1),
                  "fact mean %.2f max %d ipm %.2f pol %.2f" % (stats[:, 0].mean(), stats[:, 0].max(), stats[:, 1].mean(), stats[:, 2].mean()),
                  "oracle it %.2f act>0 %d" % (r["n_iter"].mean(), (r["n_active"] > 0).sum()), flush=True)
            if ok.any() and eu[ok].max() > 1e-4:
                b = int(np.argmax(np.where(ok, eu, 0)))
                print("  worst", b, "gpu u0", u0[b].cpu().numpy(), "oracle", r["u0"][b], "stats", stats[b], "nact", r["n_active"][b])
        except Exception:
            traceback.print_exc()
# timing
for B in (4096, 32768):
    w = wl.independent_problems(B, seed=0)
    e = Engine(batch=B, np_=7, precision="f32")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=e.dtype, device="cuda")
    xr, ur, x0 = t(w["xr"]), t(w["ur"]), t(w["x0"])
    e.reset(xr, ur); e.set_reference(xr, ur, None)
    for _ in range(5): e.solve(x0)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(20): e.solve(x0)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 20
    print("B", B, "solve %.1f us -> %.2f M solves/s" % (ms * 1e3, B / ms / 1e3), flush=True)
