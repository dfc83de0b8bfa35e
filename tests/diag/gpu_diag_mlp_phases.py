"""Diagnostic: phase timestamps (clock64 of CTA 0, thread 0) of one profiled tcgen05 MLP launch."""
import ctypes as C
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ndp_nmpc_qd_b200 import _lib
from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN

nn = DownwashNN()
lib = _lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 86016
x = torch.randn(M, 6, device="cuda")
for rep in range(3):
    y = nn.forward_rows(x, path=102)
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
lib.ndp_debug_mlp_prof(buf)
t = np.array(buf[:], dtype=np.int64)
n = int(np.argmax(t[1:] < t[:-1]) + 1) if np.any(t[1:] < t[:-1]) else int((t > 0).sum())
t = t[:n]
d = np.diff(t)
print("stamps", n, "total cycles", t[-1] - t[0], "= %.2f us at 1.965 GHz" % ((t[-1] - t[0]) / 1965.0))
print("deltas (cycles):", d.tolist())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = []
for rep in range(20):
    e0.record(); nn.forward_rows(x, path=2); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
print("kernel (events, warm L2) us:", np.median(ms) * 1e3)
