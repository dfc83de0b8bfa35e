"""BASELINE.json config 5: horizon x batch sweep, closed loop with batched dop_sim rollouts, on one GPU
(run several processes for more).  Prints one JSON line per (N, B): control steps/s, solves/s, tracking RMSE.
  python tools/closed_loop_sweep.py [--steps 250] [--max-batch 262144] [--horizons 20,40,80]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ndp_nmpc_qd_b200 import traj_gen  # noqa: E402
from ndp_nmpc_qd_b200.closed_loop import ClosedLoop  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=250)
    ap.add_argument("--max-batch", type=int, default=262144)
    ap.add_argument("--horizons", default="20,40,80")
    ap.add_argument("--precision", default="f32")
    a = ap.parse_args()
    trs = [traj_gen.plan_named("eight_high_dyn"), traj_gen.plan_named("eight_low")]
    for N in [int(x) for x in a.horizons.split(",")]:
        for B in (1, 8, 64, 512, 4096, 32768, 262144):
            if B > a.max_batch:
                continue
            rng = np.random.default_rng(B + N)
            tid = (rng.random(B) < 0.5).astype(np.int32)
            t0 = np.array([rng.uniform(0, trs[j].duration - 6.0) for j in tid])
            cl = ClosedLoop(trs, tid, t0, N=N, precision=a.precision, offset=rng.normal(size=(B, 3)) * 2.0)
            steps = a.steps if B <= 32768 else max(25, a.steps // 5)
            for _ in range(5):
                cl.step()
            torch.cuda.synchronize()
            l0 = cl.launch_count
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sq = torch.zeros((), dtype=torch.float64, device="cuda")
            for _ in range(steps):
                cl.step()
                sq += (cl.position_error() ** 2).mean()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            st = cl.engine.status().cpu().numpy()
            stats = cl.engine.stats().cpu().numpy()
            print(json.dumps(dict(N=N, batch=B, control_steps=steps, ms_per_control_step=ms / steps, solves_per_s=B * steps / (ms * 1e-3),
                                  sim_steps_per_s=2 * B * steps / (ms * 1e-3), pos_rmse_m=float(torch.sqrt(sq / steps)),
                                  status_nonzero=int((st != 0).sum()), riccati_sweeps_mean=float(stats[:, 0].mean()),
                                  launches_per_step=(cl.launch_count - l0) / steps, precision=a.precision)), flush=True)
            del cl
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
