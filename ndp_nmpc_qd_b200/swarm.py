"""Coupled swarm (BASELINE.json config 4): egos sharded over the ranks of one node, one exchange of reference
horizons per RTI step for the downwash features, nothing else crosses GPUs.

Reference semantics generalised: the reference's leader evaluates the MLP against ONE neighbour whose
reference horizon arrives as a PredXU message (ndp_nmpc_leader_node.py:60-76, nmpc_node.py:116-133); here
every ego sums the MLP force over all neighbours inside the 1 m horizontal gate.

Two exchange modes, same results (mode="auto" picks by the size of the exchanged tensor):
  "p2p"        each rank writes its [n_local, N+1, 6] fp32 horizons into a symmetric-memory buffer; after one
               device-side barrier the gating / MLP kernels read the peers' shards directly over NVLink
               (ndp_mlp_forward_swarm_parts): the transfer is fused into the feature construction.
  "allgather"  NCCL all_gather into a local [n_all, N+1, 6] buffer, then the same kernels on local memory
               (the baseline the fused mode is measured against).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .dnwash_nn_est import DownwashNN
from .params import downwash_params as DP
from .solver import Engine


def shard_bounds(n_all: int, world: int, rank: int):
    """Equal contiguous shards (the last ranks may be padded): returns (part_rows, begin, end)."""
    part = (n_all + world - 1) // world
    b = min(n_all, rank * part)
    return part, b, min(n_all, b + part)


def pack_horizons(xr: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """[n, N+1, 10] state horizons -> [n, N+1, 6] fp32 positions + velocities (the columns DownwashNN reads)."""
    out.copy_(xr[:, :, 0:6])
    return out


class SwarmStep:
    def __init__(self, n_all: int, N: int = 20, precision: str = "f32", mode: str = "p2p", device=None, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n_all, self.N = n_all, N
        self.part, self.begin, self.end = shard_bounds(n_all, self.world, self.rank)
        self.n_local = self.end - self.begin
        if mode == "auto":
            # measured on 4 / 8 B200 (profiles/r1_swarm_n*.json): peer-memory reads win while the step is latency-bound
            # (1024 quads: 174 vs 185 us per step at 8 GPUs); with several MB of horizons the repeated remote reads of the
            # MLP's feature fetch cost more than one all-gather (8192 quads: 489 vs 444 us)
            mode = "p2p" if n_all * (N + 1) * 6 * 4 <= (1 << 20) else "allgather"
        self.mode = mode if self.world > 1 else "local"
        self.dtype = torch.float32 if precision == "f32" else torch.float64
        self.engine = Engine(batch=max(self.n_local, 1), N=N, np_=7, precision=precision, device=self.device)
        self.nn = DownwashNN(device=self.device)
        self.f = torch.zeros((max(self.n_local, 1), N + 1, 3), dtype=self.dtype, device=self.device)
        self.step_no = 0
        self.trace = None  # set to a list to collect (label, cuda event) marks of every step (tests/diag/swarm_multi_gpu.py --trace)
        shape = (2, self.part, N + 1, 6)  # double-buffered by step parity: one barrier per step is enough
        if self.mode == "p2p":
            import torch.distributed._symmetric_memory as symm

            self.buf = symm.empty(shape, dtype=torch.float32, device=self.device)
            self.hdl = symm.rendezvous(self.buf, dist.group.WORLD if group is None else group)
            esz = self.part * (N + 1) * 6 * 4
            self._ptrs = [[int(p) + par * esz for p in self.hdl.buffer_ptrs] for par in (0, 1)]
        else:
            self.buf = torch.zeros(shape, dtype=torch.float32, device=self.device)
            self.gathered = torch.zeros((self.world * self.part, N + 1, 6), dtype=torch.float32, device=self.device)
        self.buf.zero_()

    def forces(self, xr_local: torch.Tensor, odom_xy: Optional[torch.Tensor] = None) -> torch.Tensor:
        """xr_local [n_local, N+1, 10]: this rank's reference horizons.  Returns f [n_local, N+1, 3]."""
        par = self.step_no & 1
        self.step_no += 1
        mine = self.buf[par]
        self._mark("begin")
        if self.n_local:
            pack_horizons(xr_local, mine[: self.n_local])
        self._mark("pack")
        if self.mode == "p2p":
            self.hdl.barrier(channel=par)  # every shard of this parity is written; peers are done with its previous use
            self._mark("exchange")
            parts, part_rows = self._ptrs[par], self.part
        elif self.mode == "allgather":
            self.dist.all_gather_into_tensor(self.gathered, mine, group=self.group)
            self._mark("exchange")
            parts, part_rows = [self.gathered.data_ptr()], self.world * self.part
        else:
            parts, part_rows = [mine.data_ptr()], self.part
        if self.n_local:
            ptrs = (C.c_void_p * len(parts))(*parts)
            _lib.check(self.nn.lib.ndp_mlp_forward_swarm_parts(
                self.nn._h, _lib.NDP_F32 if self.dtype == torch.float32 else _lib.NDP_F64, len(parts), ptrs, part_rows, self.n_all,
                self.begin, self.n_local, self.N + 1, None if odom_xy is None else C.c_void_p(odom_xy.data_ptr()), float(DP.r_horiz),
                C.c_void_p(self.f.data_ptr()), 0, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "ndp_mlp_forward_swarm_parts")
        self._mark("forces")
        return self.f

    def _mark(self, label: str):
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.trace.append((label, ev))

    def step(self, x0: torch.Tensor, xr: torch.Tensor, ur: torch.Tensor, odom_xy: Optional[torch.Tensor] = None,
             u0: Optional[torch.Tensor] = None) -> torch.Tensor:
        """one RTI step of the local egos: exchange + gated all-pairs MLP + NDP-NMPC update."""
        f = self.forces(xr, odom_xy)
        u0 = self.engine.update(x0, xr, ur, f, u0)
        self._mark("update")
        return u0


def lattice_swarm(n_all: int, seed: int = 0):
    """quads on a square lattice of 0.8 m pitch, altitude offsets U(0, 3) m, phase-shifted
    eight_low references.  Returns (position offsets [n,3], trajectory phases [n])."""
    import numpy as np

    side = int(np.ceil(np.sqrt(n_all)))
    rng = np.random.default_rng(seed)
    off = np.stack([(np.arange(n_all) % side) * 0.8, (np.arange(n_all) // side) * 0.8, rng.uniform(0.0, 3.0, n_all)], 1)
    return off, rng.uniform(0, 20.0, n_all)


def time_swarm(n_all: int, modes, steps: int = 50, warmup: int = 100, device=None, N: int = 20, use_graph: bool = True) -> dict:
    """Times the coupled-swarm RTI step (reference generation + exchange + gated all-pairs MLP + local solves) of
    `n_all` quads sharded over the ranks of the initialised process group, for each exchange mode; device time, max over
    ranks.  With more than one mode the forces of the modes are compared bit for bit."""
    import numpy as np
    import torch.distributed as dist

    from . import traj_gen
    from .traj_gen.refgen import RefGen

    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    off, t0 = lattice_swarm(n_all)
    rg = RefGen([traj_gen.plan_named("eight_low")], device=dev)
    out, forces = {}, {}
    # every mode is measured twice in turn and the second pass is reported: the first timed loop over freshly allocated
    # buffers runs 2-3 x slower than the same loop a moment later, whichever mode it is (tests/diag/swarm_multi_gpu.py)
    for rep, mode in [(r, m) for r in range(2) for m in modes]:
        sw = SwarmStep(n_all, N=N, mode=mode, device=dev)
        b, e = sw.begin, sw.end
        t_loc = torch.as_tensor(t0[b:e], device=dev)
        off_loc = torch.as_tensor(off[b:e], device=dev).contiguous()
        xr, ur = rg.horizon(t_loc, None, N, 0.1, off_loc)
        x0 = xr[:, 0].contiguous()
        sw.engine.reset(xr, ur)
        u0 = torch.empty((max(e - b, 1), 4), dtype=torch.float32, device=dev)
        def one_step():
            t_loc.add_(0.02)
            rg.horizon(t_loc, None, N, 0.1, off_loc, xr=xr, ur=ur)
            sw.step(x0, xr, ur, None, u0)

        for _ in range(warmup):
            sw.step(x0, xr, ur, None, u0)
        torch.cuda.synchronize(dev)
        # The step (clock update, reference generation, exchange, pair list + MLP + sum, the two solver kernels) is captured
        # once into a CUDA graph and replayed, like the closed loop of config 5: ten dependent launches of a few us each
        # otherwise leave a launch gap apiece.  Only when the pair buffers are sized for the worst case (the pair count then
        # never leaves the device); every rank takes the same decision.
        graph, launch_mode = None, "eager"
        if use_graph and n_all * (n_all - 1) // world <= (2 << 20):
            ok = torch.ones(1, dtype=torch.int64, device=dev)
            try:
                cap = torch.cuda.Stream(device=dev)
                cap.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(cap):
                    one_step()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=cap):
                        one_step()
                torch.cuda.current_stream(dev).wait_stream(cap)
                graph = g
            except Exception:  # noqa: BLE001
                ok.zero_()
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0:
                graph = None
            else:
                launch_mode = "cuda graph replay"
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            if graph is not None:
                graph.replay()
            else:
                one_step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        bad = torch.tensor([int((sw.engine.status()[: e - b] != 0).sum())], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(bad)
        stats = sw.engine.stats()[: e - b]
        first = out.get(sw.mode, {}).get("ms_per_step")
        out[sw.mode] = dict(ms_per_step=float(ms) / steps, quad_steps_per_s=n_all * steps / (float(ms) * 1e-3), status_nonzero=int(bad),
                            riccati_sweeps_max_rank0=int(stats[:, 0].max()) if e > b else 0, launch_mode=launch_mode)
        if first is not None:
            out[sw.mode]["ms_per_step_first_pass"] = first
        forces[sw.mode] = sw.f[: e - b].clone()
        del sw
    if len(forces) > 1:
        ks = list(forces)
        same = all(torch.equal(forces[ks[0]], forces[k]) for k in ks[1:])
        flag = torch.tensor([int(same)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["modes_bit_identical"] = bool(int(flag))
    return out
