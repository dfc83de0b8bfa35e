"""Single-GPU diagnosis of the coupled-swarm RTI step (config 4): solver statistics of the local solves
(Riccati factorisations, IPM iterations, active-set rounds, active bounds) and the time of each part.
  python tools/swarm_diag.py [--quads 1024] [--steps 20]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ndp_nmpc_qd_b200 import traj_gen  # noqa: E402
from ndp_nmpc_qd_b200.swarm import SwarmStep  # noqa: E402
from ndp_nmpc_qd_b200.traj_gen.refgen import RefGen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quads", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dense", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    n_all = a.quads
    side = int(np.ceil(np.sqrt(n_all)))
    rng = np.random.default_rng(0)
    off = np.stack([(np.arange(n_all) % side) * 0.8, (np.arange(n_all) // side) * 0.8, rng.uniform(0.0, 3.0, n_all)], 1)
    t0 = rng.uniform(0, 20.0, n_all)
    rg = RefGen([traj_gen.plan_named("eight_low")], device=dev)
    sw = SwarmStep(n_all, mode="local", device=dev)
    t_loc = torch.as_tensor(t0, device=dev)
    off_loc = torch.as_tensor(off, device=dev).contiguous()
    xr, ur = rg.horizon(t_loc, None, 20, 0.1, off_loc)
    x0 = xr[:, 0].contiguous()
    sw.engine.reset(xr, ur)
    u0 = torch.empty((n_all, 4), dtype=torch.float32, device=dev)
    out = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for s in range(a.steps):
        t_loc.add_(0.02)
        rg.horizon(t_loc, None, 20, 0.1, off_loc, xr=xr, ur=ur)
        ev[0].record()
        f = sw.forces(xr)
        ev[1].record()
        sw.engine.update(x0, xr, ur, f, u0)
        ev[2].record()
        torch.cuda.synchronize()
        st = sw.engine.stats().cpu().numpy()
        status = sw.engine.status().cpu().numpy()
        out.append(dict(step=s, forces_us=ev[0].elapsed_time(ev[1]) * 1e3, update_us=ev[1].elapsed_time(ev[2]) * 1e3,
                        constrained=int((st[:, 0] > 1).sum()), fact_mean=float(st[:, 0].mean()), fact_max=int(st[:, 0].max()),
                        ipm_max=int(st[:, 1].max()), ipm_mean_c=float(st[st[:, 0] > 1, 1].mean()) if (st[:, 0] > 1).any() else 0.0,
                        pol_max=int(st[:, 2].max()), nact_max=int(st[:, 3].max()), status_nonzero=int((status != 0).sum()),
                        f_abs_max=float(f.abs().max())))
    for o in out:
        print(json.dumps(o))
    if a.dense:
        # SURVEY.md 8(d) config 4, un-gated dense case: every pair passes the gate -> n (n - 1) x 21 MLP rows per step
        traj = xr[:, :, 0:6].float().contiguous()
        ms = []
        for s in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f = sw.nn.forward_swarm(traj, 0, n_all, r_horiz=1e6)
            e1.record()
            torch.cuda.synchronize()
            if s >= 2:
                ms.append(e0.elapsed_time(e1))
        rows = n_all * (n_all - 1) * 21
        t = float(np.mean(ms)) * 1e-3
        print(json.dumps(dict(dense_rows=rows, forces_ms=t * 1e3, mlp_rows_per_s=rows / t, algorithmic_tflops=rows * 35072 / t / 1e12,
                              f_abs_max=float(f.abs().max()))))


if __name__ == "__main__":
    main()
