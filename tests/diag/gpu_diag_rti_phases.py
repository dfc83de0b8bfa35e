"""Diagnostic (library built with -DNDP_RTI_PROF): wall-clock phases of the nominal SQP-RTI kernel, per CTA (globaltimer)."""
import ctypes as C
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ndp_nmpc_qd_b200 import _lib, workloads as wl
from ndp_nmpc_qd_b200.solver import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = wl.independent_problems(B, seed=1)
eng = Engine(batch=B, np_=7)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")
xr, ur, x0 = t(w["xr"]), t(w["ur"]), t(w["x0"])
f = torch.zeros((B, 21, 3), device="cuda")
eng.reset(xr, ur)
for _ in range(5):
    eng.update(x0, xr, ur, f)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.update(x0, xr, ur, f); e1.record(); torch.cuda.synchronize()
buf = (C.c_ulonglong * (2048 * 8))()
lib = _lib.load()
lib.ndp_debug_rti_prof(buf)
a = np.array(buf[:], dtype=np.float64).reshape(2048, 8)
full = a[a[:, 0] > 0]
a = full[:, :6]
t0 = a[:, 0].min()
print("CTAs", len(a), "event window (both kernels) us %.2f" % (e0.elapsed_time(e1) * 1e3))
for i, name in enumerate(("entry", "record staged", "cost records", "backward done", "forward done", "stored")):
    v = (a[:, i] - t0) * 1e-3
    print("%-14s min %6.2f  p10 %6.2f  median %6.2f  p90 %6.2f  max %6.2f us" % (name, v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max()))
d = np.diff(a, axis=1) * 1e-3
print("phase durations (median us):", np.round(np.median(d, axis=0), 2).tolist())

# per-SM view: how many CTAs each SM got and when its CTAs finished
sm = full[:, 6].astype(int)
end = (a[:, 5] - t0) * 1e-3
cnt = np.bincount(sm, minlength=sm.max() + 1)
print("CTAs per SM histogram:", {int(k): int((cnt == k).sum()) for k in np.unique(cnt)})
for k in np.unique(cnt):
    if k == 0:
        continue
    sel = np.isin(sm, np.nonzero(cnt == k)[0])
    print("SMs with %d CTAs: CTA finish time median %.2f  max %.2f us; backward-done median %.2f" % (k, np.median(end[sel]), end[sel].max(), np.median((a[sel, 3] - t0) * 1e-3)))
worst = np.argsort(-end)[:12]
print("slowest CTAs (block, sm, finish us):", [(int(full[i, 6] * 0 + np.nonzero(full[:, 0] > 0)[0][i]) if False else int(i), int(sm[i]), round(float(end[i]), 2)) for i in worst])
sm_end = np.array([end[sm == s_].max() if (sm == s_).any() else 0 for s_ in range(sm.max() + 1)])
print("per-SM last finish: min %.2f median %.2f max %.2f" % (sm_end[sm_end > 0].min(), np.median(sm_end[sm_end > 0]), sm_end.max()))
print("slowest SMs:", np.argsort(-sm_end)[:16].tolist(), np.round(np.sort(sm_end)[::-1][:16], 1).tolist())
print("fastest SMs:", np.argsort(sm_end + (sm_end == 0) * 1e9)[:16].tolist(), np.round(np.sort(sm_end[sm_end > 0])[:16], 1).tolist())
