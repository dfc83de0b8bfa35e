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
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of a CUDA-graph replay of the step")
    a = ap.parse_args()
    trs = [traj_gen.plan_named("eight_high_dyn"), traj_gen.plan_named("eight_low")]
    # wake the GPU up first: a batch-1 loop alone does not raise the clocks, and the first configuration would be
    # timed at idle clocks (0.55 ms instead of 0.13 ms per control step)
    warm = ClosedLoop(trs, np.zeros(4096, np.int32), np.zeros(4096), N=20, precision=a.precision)
    for _ in range(300):
        warm.step()
    torch.cuda.synchronize()
    del warm
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
            # the control step is 6 kernels of the library + one clock update: captured once into a CUDA graph and
            # replayed, so that small batches measure the device and not the Python launch rate; the tracking-error
            # diagnostics run in a second, untimed pass
            l0 = cl.launch_count
            graph, mode = None, "eager"
            if not a.eager:
                try:
                    cap = torch.cuda.Stream()
                    cap.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(cap):
                        cl.step()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=cap):
                            cl.step()
                    torch.cuda.current_stream().wait_stream(cap)
                    graph, mode = g, "cuda graph replay"
                except Exception as ex:  # noqa: BLE001
                    sys.stderr.write(f"graph capture failed ({ex}); timing eager launches\n")
            launches_per_step = float(cl.launch_count - l0) / (2 if graph is not None else 1) if graph is not None else None
            torch.cuda.synchronize()
            l0 = cl.launch_count
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                if graph is not None:
                    graph.replay()
                else:
                    cl.step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if launches_per_step is None:
                launches_per_step = (cl.launch_count - l0) / steps
            sq = torch.zeros((), dtype=torch.float64, device="cuda")
            n_err = min(steps, 100)
            for _ in range(n_err):
                cl.step()
                sq += (cl.position_error() ** 2).mean()
            torch.cuda.synchronize()
            st = cl.engine.status().cpu().numpy()
            stats = cl.engine.stats().cpu().numpy()
            print(json.dumps(dict(N=N, batch=B, control_steps=steps, ms_per_control_step=ms / steps, solves_per_s=B * steps / (ms * 1e-3),
                                  sim_steps_per_s=2 * B * steps / (ms * 1e-3), pos_rmse_m=float(torch.sqrt(sq / n_err)),
                                  status_nonzero=int((st != 0).sum()), riccati_sweeps_mean=float(stats[:, 0].mean()),
                                  launches_per_step=launches_per_step, launch_mode=mode, precision=a.precision)), flush=True)
            del cl
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
