import sys, os, ctypes as C
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
from ndp_nmpc_qd_b200 import _lib
nn = DownwashNN(); lib = _lib.load()
M = 86016
x = torch.randn(M, 6, device="cuda")
for _ in range(3): nn.forward_rows(x, path=2)
torch.cuda.synchronize()
nn.forward_rows(x, path=102); torch.cuda.synchronize()
buf = (C.c_longlong * 128)(); lib.ndp_debug_mlp_prof(buf)
ts = np.array(buf[:40]); d = np.diff(ts)
print("total 3 tiles", ts[34]-ts[0])
names = ["setup->loop"] + ["prefetch", "layer1", "sync1", "mma1 issue", "mma1 wait", "epi1", "sync2", "mma2 issue", "mma2 wait", "epi2", "store+sync"] * 3
for n, v in zip(names, d[:34]): print("%-12s %8d cyc" % (n, v))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20): nn.forward_rows(x, path=2)
ev1.record(); torch.cuda.synchronize(); print("rows kernel us", ev0.elapsed_time(ev1) / 20 * 1e3)
