#!/usr/bin/env python
"""Benchmark of the NMPC hot path (BASELINE.json metric: batched NMPC SQP-RTI solves/sec, N=20,
with downwash MLP; p50 step latency).

Workload (config 3 of BASELINE.json with the downwash MLP switched on): B independent single-quad
NDP-NMPC problems per GPU (default 4096), each with one neighbour inside the 1 m gate.  One "step" =
one pass of the hot path over the batch:
    forces = downwash MLP(fused features (other - ego)[0:6], gate)      [kernel 1, tcgen05]
    controller.update(): yref/p upload (the 42 solver.set calls) fused with one SQP-RTI step
    (RK4+sens -> Gauss-Newton -> Riccati/IPM -> full step)              [kernel 2]
Problems are independent, so N GPUs run N shards with no data-path collective (weak scaling).

  python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA engine
  python bench.py --impl reference ...                     # CPU arm: the fp64 oracle port (acados is
                                                           # not installable here), all host threads
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_HORIZON = 20
NX, NU = 10, 4
METRIC = "batched NMPC SQP-RTI solves/sec (N=20, w/ downwash MLP)"
UNIT = "solves/s"


def algorithmic_flop_per_solve(N: int, n_fact: float) -> float:
    """F_solve = N*13186 + n_it*(N*4275 + 2*N*1352 + 30*n_b), n_b = 4N + 3(N-1);
    n_it = Riccati factorisations actually performed (1 when no bound is active)."""
    n_b = 4 * N + 3 * (N - 1)
    return N * 13186.0 + n_fact * (N * 4275.0 + 2 * N * 1352.0 + 30.0 * n_b)


def compulsory_bytes_per_solve(N: int, elt: int = 4) -> float:
    """x0 + xr + ur + f + iterate read + iterate write (+ status)."""
    return elt * (10 + 10 * (N + 1) + 4 * N + 3 * (N + 1) + 2 * (14 * N + 10)) + 4


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))


def make_workload(B, seed, n_sets):
    """n_sets consecutive control steps of B problems (phases advance by 0.02 s, x0 re-perturbed)."""
    from ndp_nmpc_qd_b200 import workloads as wl

    sets = []
    base = wl.independent_problems(B, N=N_HORIZON, seed=seed, with_neighbour=True)
    rng = np.random.default_rng(seed + 1)
    for s in range(n_sets):
        w = {k: v.copy() for k, v in base.items()}
        w["x0"] = base["x0"] + 0.02 * rng.normal(size=base["x0"].shape) * np.array([1, 1, 1, 2, 2, 2, 0.2, 0.2, 0.2, 0.2])
        w["x0"][:, 6:10] /= np.linalg.norm(w["x0"][:, 6:10], axis=1, keepdims=True)
        sets.append(w)
    return sets


def cpu_oracle_rate(sets, seconds=12.0, threads=0):
    """Time the fp64 oracle port (HPIPM-like tolerance) + numpy MLP on a bounded sample, on every host core this
    process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    if threads <= 0:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    from oracle import mlp_numpy
    from oracle.c_oracle import COracle, make_cfg
    from ndp_nmpc_qd_b200.dnwash_nn_est.downwash_nn import DEFAULT_WEIGHTS

    co = COracle()
    cfg = make_cfg(tol=1e-8, max_iter=50)
    wts = mlp_numpy.load_npz(DEFAULT_WEIGHTS)
    w = sets[0]
    B = w["x0"].shape[0]
    n = min(B, 256)

    def run(n):
        X, U = w["xr"][:n].copy(), w["ur"][:n].copy()
        t0 = time.perf_counter()
        f = mlp_numpy.gated_pairs(wts, w["xr"][:n], w["other"][:n], w["xr"][:n, 0, 0:2], 1.0, np.float32).astype(np.float64)
        r = co.rti_batch(cfg, w["x0"][:n], w["xr"][:n], w["ur"][:n], f, X, U, nthreads=threads)
        return time.perf_counter() - t0, r

    run(n)  # warm up OpenMP
    dt, r = run(n)
    reps = max(1, int(np.ceil(seconds / max(dt * B / n, 1e-3))))  # whole steps of B problems until ~`seconds` of CPU work
    t_tot, n_tot, its = 0.0, 0, []
    for _ in range(reps):
        dt, r = run(B)
        t_tot += dt; n_tot += B; its.append(float(r["n_iter"].mean()))
    return dict(value=n_tot / t_tot, unit=UNIT, cores=int(r["threads"]), kind="port",
                sample=f"{reps} x the {B} problems of one step ({n_tot} solves), fp64 C restatement of acados SQP_RTI+HPIPM (tol 1e-8) + numpy MLP, {t_tot:.1f} s",
                n_iter_mean=float(np.mean(its))), n_tot, t_tot


def acados_probe():
    """look for a real acados before falling back to the oracle port -- `import acados_template`,
    $ACADOS_SOURCE_DIR, baseline/_ref/.  Reports what it found; this image has none of them."""
    found = {}
    try:
        import acados_template  # noqa: F401

        found["acados_template"] = getattr(acados_template, "__file__", "importable")
    except Exception as e:
        found["acados_template"] = f"not importable ({type(e).__name__})"
    try:
        import casadi  # noqa: F401

        found["casadi"] = "importable"
    except Exception as e:
        found["casadi"] = f"not importable ({type(e).__name__})"
    src = os.environ.get("ACADOS_SOURCE_DIR")
    found["ACADOS_SOURCE_DIR"] = src if src and os.path.isdir(src) else ("unset" if not src else f"{src} (missing)")
    ref = os.path.join(ROOT, "baseline", "_ref")
    found["baseline/_ref"] = sorted(os.listdir(ref))[:8] if os.path.isdir(ref) else "absent"
    found["usable"] = bool(found["acados_template"].startswith("/") and found["casadi"] == "importable" and src and os.path.isdir(src))
    return found


def run_reference(args, rank, world):
    if rank != 0:
        return
    probe = acados_probe()
    sets = make_workload(args.batch, 0, 1)
    times = []
    info = None
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for s in range(args.warmup + args.steps):
        info, n, dt = cpu_oracle_rate(sets, seconds=per_step)
        if s >= args.warmup:
            times.append((n, dt))
    tot_n, tot_t = sum(n for n, _ in times), sum(t for _, t in times)
    value = tot_n / tot_t
    info["value"] = value
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=1e3 * tot_t / len(times), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
               data="synthetic", impl="reference",
               config=dict(workload=f"config 3: {args.batch} independent single-quad NDP-NMPC problems per GPU, N=20, one gated neighbour each (downwash MLP on)",
                           note="acados/HPIPM/CasADi are not installable here (un-vendored, no network); this arm is the fp64 CPU oracle port on all host threads; each step is a bounded sample"),
               cpu_baseline=info, e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
               acados_probe=probe)
    print(json.dumps(out), flush=True)


def run_native(args, rank, local_rank, world):
    import torch

    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.solver import Engine

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, K, W = args.batch, args.steps, args.warmup
    n_sets = 8
    sets = make_workload(B, seed=1000 * rank, n_sets=n_sets)
    f64 = args.dtype == "f64"
    dt = torch.float64 if f64 else torch.float32
    eng = Engine(batch=B, N=N_HORIZON, np_=7, precision=args.dtype, device=dev)
    nn = DownwashNN(device=dev)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    d_sets = [dict(x0=t(w["x0"]), xr=t(w["xr"]), ur=t(w["ur"]), other=t(w["other"]), gate=t(w["xr"][:, 0, 0:2])) for w in sets]
    f_buf = torch.empty((B, N_HORIZON + 1, 3), dtype=dt, device=dev)
    u0_buf = torch.empty((B, NU), dtype=dt, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step(d):
        nn.forward_pairs(d["xr"], d["other"], d["gate"], out=f_buf)
        eng.update(d["x0"], d["xr"], d["ur"], f_buf, u0_buf, f_from_prev_kernel=True)

    eng.reset(d_sets[0]["xr"], d_sets[0]["ur"])
    for s in range(W):
        step(d_sets[s % n_sets])
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    # ---------- device-resident timing (value) ----------
    # the step as a user runs it: MLP kernel, then the solve as its programmatic dependent (no event between the two)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    l0 = eng.launch_count + nn.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize()
    for s in range(K):
        d = d_sets[(W + s) % n_sets]
        flush.zero_()  # evict L2 between timed iterations
        ev[s][0].record()
        step(d)
        ev[s][1].record()
    torch.cuda.synchronize()
    launches = eng.launch_count + nn.launch_count - l0
    step_ms = np.array([e[0].elapsed_time(e[1]) for e in ev])
    # per-kernel durations for the roofline: the same K steps again with an event between the two kernels (which
    # serialises them: an ordinary launch of the solve)
    evk = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    eng.kernel_timing(True)  # the library brackets its two kernels (nominal SQP-RTI, constrained QPs) with events of its own
    nom_ms, con_ms = [], []
    for s in range(K):
        d = d_sets[(W + s) % n_sets]
        flush.zero_()
        evk[s][0].record()
        nn.forward_pairs(d["xr"], d["other"], d["gate"], out=f_buf)
        evk[s][1].record()
        eng.update(d["x0"], d["xr"], d["ur"], f_buf, u0_buf)
        evk[s][2].record()
        a_ms, c_ms = eng.last_kernel_ms()
        nom_ms.append(a_ms); con_ms.append(c_ms)
    torch.cuda.synchronize()
    eng.kernel_timing(False)
    solve_ms = np.array(nom_ms)  # the dominant kernel's own duration (roofline); the constrained kernel is listed beside it
    mlp_ms = np.array([e[0].elapsed_time(e[1]) for e in evk])
    serial_step_ms = np.array([e[0].elapsed_time(e[2]) for e in evk])
    total_ms = float(step_ms.sum())
    stats = eng.stats().cpu().numpy()
    status = eng.status().cpu().numpy()
    if args.kernels_only:
        if rank == 0:
            print(json.dumps(dict(kernels_only=True, ms_per_step=total_ms / K, rti_ms=float(solve_ms.mean()), mlp_ms=float(mlp_ms.mean()))), flush=True)
        return
    # ---------- end-to-end timing through the host-buffer C ABI (e2e) ----------
    # ndp_pipeline_* in long-list mode: the reference lists live on the device (as in the reference's own publisher, which
    # appends ONE new point per tick: pt_publisher.py:78-97) and every step copies that step's record -- x0, the new ego
    # point, the neighbour's new point, the gate position: 128 B per problem -- from pinned host memory, runs the list push,
    # the MLP and the RTI kernels and copies (u0, status) back; consecutive steps overlap (upload of step i+1 under the
    # kernels of step i), results are consumed on the host in order.  Every step has its own pre-filled pinned slot.
    from ndp_nmpc_qd_b200 import workloads as wl
    from ndp_nmpc_qd_b200.pipeline import HostStepPipeline, LongList

    inflight = 4
    n_lat = min(W + min(K, 30), 60)            # one-step-at-a-time latency samples (the first W are warm-up)
    n_thr = min(K, 448 - n_lat)                # overlapped steps timed for the throughput
    n_slots = n_lat + n_thr
    lists = wl.sliding_lists(sets[0], n_slots + 1)
    ll = LongList(eng, with_other=True)
    pipe = HostStepPipeline(eng, nn, depth=n_slots, longlist=ll)
    rng = np.random.default_rng(7 + rank)
    for j, sl in enumerate(pipe.slots):        # tick j + 1: the lists hold points j + 1 .. j + 101
        x0 = lists["x_list"][:, j + 1] + 0.02 * rng.normal(size=(B, 10)) * np.array([1, 1, 1, 2, 2, 2, 0.2, 0.2, 0.2, 0.2])
        x0[:, 6:10] /= np.linalg.norm(x0[:, 6:10], axis=1, keepdims=True)
        sl.x0[...] = x0; sl.xr[:, 0] = lists["x_list"][:, j + 101]; sl.ur[:, 0] = lists["u_list"][:, j + 101]
        sl.other[:, 0] = lists["other_list"][:, j + 101]; sl.gate_xy[...] = x0[:, 0:2]
    ll.reset(lists["x_list"][:, :101], lists["u_list"][:, :101], lists["other_list"][:, :101])
    eng.reset(d_sets[0]["xr"], d_sets[0]["ur"])
    torch.cuda.synchronize()
    # latency: one step at a time (submit + wait)
    e2e_lat = []
    for s in range(n_lat):
        t0 = time.perf_counter()
        pipe.step(s)
        if s >= W:
            e2e_lat.append(time.perf_counter() - t0)
    chk = float(pipe.slots[0].u0[0, 3])
    assert chk == chk
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    # throughput: `inflight` steps in flight
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cs = torch.cuda.ExternalStream(pipe.lib.ndp_pipeline_stream(pipe._p), device=dev)
    t_host0 = time.perf_counter()
    e0.record(cs)
    acc = 0.0
    for s in range(n_lat, n_slots):
        if s - n_lat >= inflight:
            acc += float(pipe.wait(s - inflight).u0[0, 3])  # the host reads a step's result while later steps are in flight
        pipe.submit(s)
    for s in range(max(n_lat, n_slots - inflight), n_slots):
        acc += float(pipe.wait(s).u0[0, 3])
    e1.record(cs)
    torch.cuda.synchronize()
    t_host = (time.perf_counter() - t_host0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), t_host)  # device span of the steps vs host wall clock incl. the last D2H: report the slower
    e2e_bad = int(sum(int((pipe.slots[s].status != 0).sum()) for s in range(n_lat, n_slots)))
    h2d_b, d2h_b = pipe.h2d_bytes_per_step, pipe.d2h_bytes_per_step
    e2e_steps = n_thr
    depth = inflight
    clocks = sampler.stop() if sampler else None
    # ---------- configs 4 and 5 under the same clock (all ranks take part) ----------
    extras = None if args.no_extras else config45(args, rank, world, dev, dist)
    # ---------- reduce over ranks: max time ----------
    red = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.barrier()
    total_ms, e2e_ms = float(red[0]), float(red[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    n_fact = float(stats[:, 0].mean())
    flop = algorithmic_flop_per_solve(N_HORIZON, n_fact)
    elt = 8 if f64 else 4
    lanes = 64 if f64 else 128  # FMA lanes per SM of the pipe the kernel computes in (FP64: 64 per SM, spec, unmeasured)
    fp32_peak = 148 * lanes * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12  # TFLOP/s, CUDA-core FMA pipe
    solve_s = float(solve_ms.mean()) * 1e-3
    ach_tf = flop * B / solve_s / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_rti_summary_f64.json" if f64 else "ncu_rti_summary.json")
    traffic_src = None
    co_bounds = None
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            traffic = pj.get("dram_bytes_per_launch")
            r = pj.get("rti", {})
            if r.get("smem_wavefronts_per_launch") and r.get("warp_instructions_per_launch"):
                # what the kernel actually runs out of: shared-memory wavefronts (1 per clock and SM) and issue
                # slots (1 per clock and scheduler), counted under ncu, over this run's own kernel time
                clk = peaks["sm_max_mhz"] * 1e6
                co_bounds = dict(smem_wavefronts_per_launch=r["smem_wavefronts_per_launch"],
                                 smem_frac_of_peak=r["smem_wavefronts_per_launch"] / (148 * clk * solve_s),
                                 warp_instructions_per_launch=r["warp_instructions_per_launch"],
                                 issue_frac_of_peak=r["warp_instructions_per_launch"] / (148 * 4 * clk * solve_s),
                                 note="the kernel is co-bound by shared-memory wavefronts and issue slots (4096 problems are 3.46 warps per scheduler: "
                                      "the schedulers holding 4 warps, ~75 % issue-active, set the launch time)")
            traffic_src = f"{os.path.relpath(prof, ROOT)} written by `bench.py --profile` on {pj.get('when', pj.get('source', '?'))}"
        except Exception:
            traffic = None
    out = dict(
        metric=METRIC, value=world * B * K / (total_ms * 1e-3), unit=UNIT, n_gpus=world, steps=K, warmup=W,
        ms_per_step=total_ms / K, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype, data="synthetic",
        config=dict(workload=f"config 3: {B} independent single-quad NDP-NMPC problems per GPU, N=20, one gated neighbour each (downwash MLP on)",
                    batch_per_gpu=B, horizon=N_HORIZON, l2="flushed between timed iterations (256 MiB memset outside the event pair)",
                    inputs="8 pre-generated control steps cycled; iterate warm-started, no shift",
                    parallelism=f"{world} independent shards, no collective"),
        clocks=clocks,
        e2e=dict(value=world * B * e2e_steps / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=int(h2d_b), d2h_bytes_per_step=int(d2h_b),
                 ms_per_step=e2e_ms / e2e_steps, steps=e2e_steps, p50_step_ms=float(np.median(e2e_lat) * 1e3), status_nonzero=e2e_bad,
                 api="ndp_pipeline_create_ll / submit / wait (include/ndp_nmpc.h): pinned host record -> H2D -> list push + MLP + RTI kernels -> D2H (u0, status)",
                 inputs="per step and problem: x0, ONE new reference point (x, u), the neighbour's new point, the gate position (128 B); "
                        "the 101-point reference lists live on the device, as the reference's publisher keeps them (pt_publisher.py:78-97)",
                 mode=f"{depth} steps in flight (upload of step i+1 overlaps the kernels of step i); p50_step_ms is the one-step-at-a-time latency"),
        gpu_launches=int(launches),
        roofline=dict(bound="fp64" if f64 else "fp32", kernel="rti_step_kernel<double>" if f64 else "rti_step_kernel<float>", achieved=ach_tf, peak=fp32_peak,
                      unit="TFLOP/s", frac=ach_tf / fp32_peak,
                      traffic=traffic, traffic_source=traffic_src, co_bounds=co_bounds, flop_per_solve=flop, n_fact_mean=n_fact, kernel_ms=float(solve_ms.mean()),
                      peak_source=f"148 SM x {lanes} FMA x 2 x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} sm_max_mhz)",
                      hbm=dict(achieved=compulsory_bytes_per_solve(N_HORIZON, elt) * B / solve_s / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                               frac=compulsory_bytes_per_solve(N_HORIZON, elt) * B / solve_s / 1e9 / peaks["hbm_gbs"],
                               bytes_per_solve=compulsory_bytes_per_solve(N_HORIZON, elt))),
        mlp=dict(kernel_ms=float(mlp_ms.mean()), rows=B * (N_HORIZON + 1), note="fused feature + gate + MLP kernel (events 0-1 of the per-kernel pass)"),
        constrained_kernel=dict(kernel_ms=float(np.mean(con_ms)), note="rti_constrained_kernel of the same steps: takes the problems whose unconstrained step leaves its box (none on this workload: it reads an empty queue and exits)"),
        step_breakdown=dict(serialised_step_ms=float(serial_step_ms.mean()),
                            note="value times the step with the solve launched as a programmatic dependent of the MLP kernel; "
                                 "roofline.kernel_ms / mlp.kernel_ms come from a second pass of the same steps with an event between the two kernels"),
        p50_step_ms=float(np.median(step_ms)), p99_step_ms=float(np.quantile(step_ms, 0.99)),
        solver=dict(status_nonzero=int((status != 0).sum()), ipm_iters_mean=float(stats[:, 1].mean()), active_bounds_mean=float(stats[:, 3].mean())),
    )
    # parity of the benched problems against the CPU oracle (a parity figure with every number)
    out["parity"] = parity_check(eng, nn, sets, dev, dt, f_buf, u0_buf)
    if extras:
        out.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_oracle_rate(sets, seconds=12.0)[0]
    if world == 1 and not args.no_latency and not f64:
        out["latency_b1"] = batch1_latency(dev)
        out["stress"] = stress_variant(dev, B)
        # fp32 build's first step on the same problems, for the fp32-vs-fp64 distance
        eng.reset(d_sets[0]["xr"], d_sets[0]["ur"])
        nn.forward_pairs(d_sets[0]["xr"], d_sets[0]["other"], d_sets[0]["gate"], out=f_buf)
        eng.update(d_sets[0]["x0"], d_sets[0]["xr"], d_sets[0]["ur"], f_buf, u0_buf)
        torch.cuda.synchronize()
        out["f64_build"] = f64_build(dev, sets[0], u0_buf.cpu().numpy().astype(np.float64))
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def config45(args, rank, world, dev, dist):
    """BASELINE.json config 4 (coupled swarm: the one case with an exchange step) and config 5 (horizon x batch closed-loop
    sweep with batched dop_sim rollouts) on the ranks of this run.
      swarm        1024 (and 8192) quads in total, sharded over the ranks (STRONG scaling: the step is latency-bound),
                   per exchange mode: one process -> local; several -> fused peer-memory reads (p2p) and NCCL all-gather
      closed_loop  every rank runs its own scenarios (weak scaling, no collective): ms per control step, max over ranks"""
    import torch

    from ndp_nmpc_qd_b200.closed_loop import time_closed_loop
    from ndp_nmpc_qd_b200.swarm import time_swarm

    out = {}
    modes = ["local"] if world == 1 else ["p2p", "allgather"]
    sw = {}
    for quads in (1024, 8192):
        try:
            sw[str(quads)] = time_swarm(quads, modes, steps=40, warmup=60, device=dev)
        except Exception as ex:  # noqa: BLE001 -- e.g. symmetric memory unavailable on this box
            sw[str(quads)] = dict(error=f"{type(ex).__name__}: {ex}"[:300])
    out["swarm"] = dict(nranks=world, scaling="strong", quads=sw,
                        note="ms per coupled RTI step: reference generation + exchange + gated all-pairs MLP + local solves; device time, max over ranks")
    cl = []
    for N, B, steps in ((20, 32768, 60), (40, 32768, 40), (80, 32768, 30), (20, 262144, 20), (80, 262144, 10)):
        try:
            r = time_closed_loop(N, B, steps=steps, device=dev, seed=1000 * rank + N + B)
            ms = torch.tensor([r["ms_per_control_step"]], dtype=torch.float64, device=dev)
            bad = torch.tensor([r["status_nonzero"]], dtype=torch.int64, device=dev)
            if dist is not None:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(bad)
            cl.append(dict(N=N, batch_per_gpu=B, control_steps=steps, ms_per_control_step=float(ms), solves_per_s=world * B / (float(ms) * 1e-3),
                           sim_steps_per_s=2 * world * B / (float(ms) * 1e-3), pos_rmse_m_rank0=r["pos_rmse_m"], status_nonzero=int(bad),
                           launch_mode=r["launch_mode"]))
        except Exception as ex:  # noqa: BLE001
            cl.append(dict(N=N, batch_per_gpu=B, error=f"{type(ex).__name__}: {ex}"[:300]))
    out["closed_loop"] = dict(nranks=world, scaling="weak", points=cl,
                              note="RefGen -> odometry -> SQP-RTI -> AttitudeTarget -> 2 plant steps per control step, CUDA-graph replay; th_pred stays 0.1 s (T = 0.1 N)")
    return out


def parity_check(eng, nn, sets, dev, dt, f_buf, u0_buf, steps=3):
    """The benched batch, `steps` warm-started control steps from the iterate reset to the reference: downwash MLP +
    SQP-RTI step on the GPU against the CPU oracle chain (numpy MLP -> fp64 C SQP-RTI oracle solved to the exact QP
    solution).  Per-component errors |a - b| / max(|b|, 1); the north-star gate is 1e-4 (fp32 build)."""
    import torch

    from oracle import mlp_numpy
    from oracle.c_oracle import COracle, make_cfg
    from ndp_nmpc_qd_b200.dnwash_nn_est.downwash_nn import DEFAULT_WEIGHTS

    co, cfg, wts = COracle(), make_cfg(), mlp_numpy.load_npz(DEFAULT_WEIGHTS)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
    w0 = sets[0]
    xr, ur = t(w0["xr"]), t(w0["ur"])
    eng.reset(xr, ur)
    X, U = w0["xr"].copy(), w0["ur"].copy()
    worst = dict(u0_rel_max=0.0, X_rel_max=0.0, U_rel_max=0.0, mlp_abs_max=0.0)
    n_ok = 0
    for s in range(steps):
        w = sets[s % len(sets)]
        nn.forward_pairs(t(w["xr"]), t(w["other"]), t(w["xr"][:, 0, 0:2]), out=f_buf)
        eng.update(t(w["x0"]), t(w["xr"]), t(w["ur"]), f_buf, u0_buf)
        torch.cuda.synchronize()
        f_ref = mlp_numpy.gated_pairs(wts, w["xr"], w["other"], w["xr"][:, 0, 0:2], 1.0, np.float32).astype(np.float64)
        r = co.rti_batch(cfg, w["x0"], w["xr"], w["ur"], f_ref, X, U)
        ok = r["status"] == 0
        n_ok += int(ok.sum())
        gX, gU = eng.get_all("x").cpu().numpy().astype(np.float64), eng.get_all("u").cpu().numpy().astype(np.float64)
        worst["u0_rel_max"] = max(worst["u0_rel_max"], rel(u0_buf.cpu().numpy().astype(np.float64)[ok], r["u0"][ok]))
        worst["X_rel_max"] = max(worst["X_rel_max"], rel(gX[ok], X[ok]))
        worst["U_rel_max"] = max(worst["U_rel_max"], rel(gU[ok], U[ok]))
        worst["mlp_abs_max"] = max(worst["mlp_abs_max"], float(np.abs(f_buf.cpu().numpy().astype(np.float64) - f_ref).max()))
    worst.update(problems=int(w0["x0"].shape[0]), steps=steps, compared=n_ok, gate=1e-4 if dt == torch.float32 else 1e-9,
                 oracle="numpy MLP + fp64 C SQP-RTI oracle (IPM + exact active-set polish), per-component |a-b|/max(|b|,1)")
    return worst


def stress_variant(dev, B, steps=20):
    """SURVEY.md 8(d) stress variant of config 3 (exercises the inequality path, which the nominal
    distribution never does): 5x perturbation scale, tightened bounds (omega_max 1.5 rad/s, c_max 15 m/s^2),
    disturbance forces ~ N(0, 1 N).  Every timed solve starts from the iterate reset to the reference."""
    import torch

    from ndp_nmpc_qd_b200 import workloads as wl
    from ndp_nmpc_qd_b200.solver import Engine

    w = wl.independent_problems(B, N=N_HORIZON, seed=5, scale=5.0)
    fd = np.random.default_rng(6).normal(size=(B, N_HORIZON + 1, 3))
    eng = Engine(batch=B, N=N_HORIZON, np_=7, precision="f32", device=dev, u_min=[-1.5, -1.5, -1.5, 0.0], u_max=[1.5, 1.5, 1.5, 15.0])
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    x0, xr, ur, f = t(w["x0"]), t(w["xr"]), t(w["ur"]), t(fd)
    u0 = torch.empty((B, NU), dtype=torch.float32, device=dev)
    ms = []
    for s in range(steps + 3):
        eng.reset(xr, ur)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.update(x0, xr, ur, f, u0)
        e1.record()
        torch.cuda.synchronize()
        if s >= 3:
            ms.append(e0.elapsed_time(e1))
    st, status = eng.stats().cpu().numpy(), eng.status().cpu().numpy()
    n_fact = float(st[:, 0].mean())
    k_ms = float(np.mean(ms))
    return dict(value=B / (k_ms * 1e-3), unit=UNIT, kernel_ms=k_ms, riccati_sweeps_mean=n_fact, riccati_sweeps_max=int(st[:, 0].max()),
                constrained_share=float((st[:, 0] > 1).mean()), ipm_share=float((st[:, 1] > 0).mean()),
                active_bounds_mean=float(st[:, 3].mean()), status_nonzero=int((status != 0).sum()),
                achieved_tflops=algorithmic_flop_per_solve(N_HORIZON, n_fact) * B / (k_ms * 1e-3) / 1e12,
                note="RTI kernel only (no MLP): 5x perturbation, omega_max 1.5, c_max 15, f ~ N(0,1); iterate reset before every solve")


def f64_build(dev, w, u0_f32):
    """The fp64 build of the same solve on one step's problems (north star: 'tighter for an fp64 build'): kernel time
    and the distance of the fp32 build's u0 from it."""
    import torch

    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.solver import Engine

    B = w["x0"].shape[0]
    eng = Engine(batch=B, N=N_HORIZON, np_=7, precision="f64", device=dev)
    nn = DownwashNN(device=dev)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    x0, xr, ur, other = t(w["x0"]), t(w["xr"]), t(w["ur"]), t(w["other"])
    f = nn.forward_pairs(xr, other, t(w["xr"][:, 0, 0:2]))
    u0 = torch.empty((B, NU), dtype=torch.float64, device=dev)
    ms = []
    for s in range(8):
        eng.reset(xr, ur)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.update(x0, xr, ur, f, u0)
        e1.record()
        torch.cuda.synchronize()
        if s >= 2:
            ms.append(e0.elapsed_time(e1))
    k_ms = float(np.mean(ms))
    ref = u0.cpu().numpy()
    rel = float((np.abs(u0_f32 - ref).max(1) / np.maximum(np.abs(ref).max(1), 1.0)).max()) if u0_f32 is not None else None
    return dict(kernel_ms=k_ms, value=B / (k_ms * 1e-3), unit=UNIT, u0_rel_diff_f32_vs_f64=rel,
                status_nonzero=int((eng.status().cpu().numpy() != 0).sum()),
                note="rti_step_kernel<double>, same problems, first RTI step from the iterate reset to the reference (MLP forces from the fp32-accurate kernel)")


def batch1_latency(dev):
    """p50 of controller.update() through the drop-in Python surface at batch 1 (configs 1/2)."""
    from ndp_nmpc_qd_b200 import workloads as wl
    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.ndp_nmpc_ctl import NDPNMPCBodyRateController

    ctl = NDPNMPCBodyRateController(device=dev)
    nn = DownwashNN(device=dev)
    xr, ur = wl.reference_horizon([1.0])
    ctl.reset(xr[0], ur[0])
    lat_u, lat_n = [], []
    for i in range(220):
        xr, ur = wl.reference_horizon([1.0 + 0.02 * i])
        other = xr[0].copy(); other[:, 2] += 0.8  # a neighbour flying 0.8 m above the ego
        t0 = time.perf_counter()
        f = nn.update(other, xr[0])
        t1 = time.perf_counter()
        u_last = ctl.update(xr[0, 0], xr[0], ur[0], f)
        t2 = time.perf_counter()
        if i >= 20:
            lat_n.append(t1 - t0); lat_u.append(t2 - t1)
    # the leader's whole tick (DownwashNN.update + controller.update) as ONE call through the C ABI:
    # ndp_pipeline_submit/wait at batch 1 (pinned host record -> H2D -> MLP + RTI kernels -> D2H)
    from ndp_nmpc_qd_b200.pipeline import HostStepPipeline
    from ndp_nmpc_qd_b200.solver import Engine
    import torch

    eng = Engine(batch=1, N=N_HORIZON, np_=7, precision="f32", device=dev)
    pipe = HostStepPipeline(eng, nn, depth=1)
    xr, ur = wl.reference_horizon([1.0])
    t32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    eng.reset(t32(xr), t32(ur))
    torch.cuda.synchronize()
    sl, lat_f, u_f = pipe.slots[0], [], None
    for i in range(220):
        xr, ur = wl.reference_horizon([1.0 + 0.02 * i])
        other = xr[0].copy(); other[:, 2] += 0.8
        t0 = time.perf_counter()
        sl.x0[...] = xr[:, 0]; sl.xr[...] = xr; sl.ur[...] = ur
        sl.other[...] = other[None, :, 0:6]; sl.gate_xy[...] = xr[:, 0, 0:2]
        u_f = pipe.step(0).u0[0].copy()
        t1 = time.perf_counter()
        if i >= 20:
            lat_f.append(t1 - t0)
    return dict(update_p50_us=float(np.median(lat_u) * 1e6), update_p99_us=float(np.quantile(lat_u, 0.99) * 1e6),
                downwash_update_p50_us=float(np.median(lat_n) * 1e6),
                fused_tick_p50_us=float(np.median(lat_f) * 1e6), fused_tick_p99_us=float(np.quantile(lat_f, 0.99) * 1e6),
                fused_tick_u0_vs_dropin=float(np.abs(u_f - u_last).max()),
                note="NDPNMPCBodyRateController.update / DownwashNN.update, host numpy in -> out; fused_tick = both in one "
                     "ndp_pipeline_submit/wait call at batch 1 (same inputs, same closed sequence)")


def run_profile(args):
    """`bench.py --profile`: regenerate the measured DRAM traffic of the dominant kernel (roofline.traffic) instead of
    trusting a committed number -- one ncu pass (dram__bytes_read/write, gpu__time_duration; --clock-control none) over a
    short kernels-only run of this same benchmark; the summary goes to profiles/ncu_rti_summary.json, which the normal
    run reads.  Numbers printed by the profiled child are never bench values."""
    import csv
    import datetime
    import tempfile

    tmp = tempfile.NamedTemporaryFile(suffix=".csv", delete=False).name
    child = [sys.executable, os.path.abspath(__file__), "--steps", "6", "--warmup", "3", "--batch", str(args.batch), "--dtype", args.dtype,
             "--kernels-only"]
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,"
           "smsp__inst_executed.sum", "--clock-control", "none",
           "-k", "regex:rti_step_kernel|mlp_tc_kernel", "-c", "80", "--csv", "--log-file", tmp] + child
    rc = subprocess.call(cmd, stdout=subprocess.DEVNULL)
    rows, header = [], None
    for row in csv.reader(open(tmp, errors="replace")):
        if header is None:
            if row and row[0] == "ID":
                header = row
            continue
        if len(row) == len(header):
            rows.append(dict(zip(header, row)))
    per = {}
    for r in rows:
        k = "rti" if "rti_step_kernel" in r["Kernel Name"] else "mlp"
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        if "byte" in unit.lower():
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        per.setdefault(k, {}).setdefault(r["ID"], {})[r["Metric Name"]] = v
    out = dict(command=" ".join(cmd), when=datetime.datetime.utcnow().isoformat() + "Z", ncu_exit_code=rc, batch=args.batch, dtype=args.dtype)
    for k, launches in per.items():
        ls = list(launches.values())[len(launches) // 2:]  # second half: past the warm-up steps
        if not ls:
            continue
        tr = [l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in ls]
        out[k] = dict(launches=len(ls), dram_bytes_per_launch=float(np.mean(tr)), dram_read=float(np.mean([l.get("dram__bytes_read.sum", 0.0) for l in ls])),
                      dram_write=float(np.mean([l.get("dram__bytes_write.sum", 0.0) for l in ls])),
                      smem_wavefronts_per_launch=float(np.mean([l.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0.0) for l in ls])),
                      warp_instructions_per_launch=float(np.mean([l.get("smsp__inst_executed.sum", 0.0) for l in ls])),
                      kernel_us_under_ncu=float(np.mean([l.get("gpu__time_duration.sum", 0.0) for l in ls])))
    if "rti" in out:
        out["dram_bytes_per_launch"] = out["rti"]["dram_bytes_per_launch"]
        out["algorithmic_bytes_per_launch"] = compulsory_bytes_per_solve(N_HORIZON, 8 if args.dtype == "f64" else 4) * args.batch
    dst = os.path.join(ROOT, "profiles", "ncu_rti_summary.json" if args.dtype == "f32" else "ncu_rti_summary_f64.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(dict(profile=dst, **{k: out[k] for k in ("rti", "mlp") if k in out})), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="engine precision (f64: the like-for-like build against the fp64 reference path)")
    ap.add_argument("--profile", action="store_true", help="re-measure roofline.traffic with ncu (writes profiles/ncu_rti_summary.json)")
    ap.add_argument("--kernels-only", action="store_true", help="device-resident timing only (what --profile runs under ncu)")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 4 (swarm) / config 5 (closed-loop sweep) blocks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.profile:
        run_profile(args)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
