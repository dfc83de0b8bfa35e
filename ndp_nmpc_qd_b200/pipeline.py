"""Host-buffer step pipeline: one control step for the whole batch straight from pinned HOST memory.

What the reference's ROS node does per 50 Hz tick -- `DownwashNN.update(other, ego_ref)`
(ndp_nmpc_leader_node.py:60-76; downwash_nn.py:21-29 incl. its H2D/D2H) then
`controller.update(x0, xr, ur, f)` (nmpc_node.py:202-209; ndp_nmpc_body_rate_ctl.py:91-112) -- as one
asynchronous submission through the C ABI (`ndp_pipeline_*`, include/ndp_nmpc.h): H2D copy of the
step record, MLP + RTI kernels, D2H copy of (u0, status).  The caller writes a slot's numpy views in
place (they alias pinned memory owned by the library), submits the slot and later waits for it;
uploads, kernels and downloads of different slots overlap.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .params import downwash_params as DP
from .solver import Engine

NX, NU = 10, 4


class LongList:
    """Device-resident sliding reference lists (ndp_longlist_*): NMPCRefPublisher's 101-point lists
    (pt_pub/pt_publisher.py:57-103) for B quadrotors, plus -- with_other -- their neighbours' lists for DownwashNN."""

    def __init__(self, engine: Engine, with_other: bool = False, stride: int = 5, length: int = 101):
        import torch

        self.engine, self.lib = engine, engine.lib
        self.B, self.N, self.stride, self.length, self.with_other = engine.batch, engine.N, int(stride), int(length), bool(with_other)
        self._l = C.c_void_p()
        with torch.cuda.device(engine.device):
            _lib.check(self.lib.ndp_longlist_create(engine.cfg.precision, self.B, self.N, self.stride, self.length, int(with_other), C.byref(self._l)),
                       "ndp_longlist_create")

    def reset(self, x_long, u_long, other_long=None, stream=None):
        """x_long [B, len, 10], u_long [B, len, 4], other_long [B, len, 6]: numpy (host) or CUDA tensors in engine precision."""
        import torch

        np_dt = np.float32 if self.engine.cfg.precision == _lib.NDP_F32 else np.float64
        keep = []

        def ptr(a, shape):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                assert a.dtype == self.engine.dtype and a.is_contiguous() and tuple(a.shape) == shape
                return C.c_void_p(a.data_ptr())
            a = np.ascontiguousarray(a, dtype=np_dt).reshape(shape)
            keep.append(a)
            return C.c_void_p(a.ctypes.data)

        s = stream if stream is not None else torch.cuda.current_stream(self.engine.device)
        _lib.check(self.lib.ndp_longlist_reset(self._l, ptr(x_long, (self.B, self.length, NX)), ptr(u_long, (self.B, self.length, NU)),
                                               ptr(other_long, (self.B, self.length, 6)), C.c_void_p(s.cuda_stream)), "ndp_longlist_reset")
        s.synchronize()  # pageable host sources must outlive the copy

    def push(self, new_x, new_u, new_other=None, xr=None, ur=None, other=None, stream=None):
        """get_nmpc_pts for the whole batch: CUDA tensors new_x [B,10], new_u [B,4], new_other [B,6] -> (xr, ur, other)."""
        import torch

        e = self.engine
        if xr is None:
            xr = torch.empty((self.B, self.N + 1, NX), dtype=e.dtype, device=e.device)
        if ur is None:
            ur = torch.empty((self.B, self.N, NU), dtype=e.dtype, device=e.device)
        if self.with_other and other is None:
            other = torch.empty((self.B, self.N + 1, 6), dtype=e.dtype, device=e.device)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        s = stream if stream is not None else torch.cuda.current_stream(e.device)
        _lib.check(self.lib.ndp_longlist_push(self._l, p(new_x), p(new_u), p(new_other), p(xr), p(ur), p(other), C.c_void_p(s.cuda_stream)),
                   "ndp_longlist_push")
        return xr, ur, other

    def close(self):
        if getattr(self, "_l", None) is not None and self._l:
            self.lib.ndp_longlist_destroy(self._l)
            self._l = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StepSlot:
    """numpy views of one slot's pinned host arrays."""
    __slots__ = ("x0", "xr", "ur", "other", "gate_xy", "u0", "status")


class HostStepPipeline:
    def __init__(self, engine: Engine, downwash=None, depth: int = 2, r_horiz: float = DP.r_horiz, longlist: Optional[LongList] = None):
        """engine: `Engine` (np_=7 when `downwash` is given); downwash: `DownwashNN` or None.
        longlist: the reference horizons live on the device (`LongList`) and a slot carries ONE new point per list --
        slot.xr is then [B, 1, 10] (the new ego point), slot.ur [B, 1, 4], slot.other [B, 1, 6]: what the reference's
        node itself produces per tick (get_nmpc_pts, pt_publisher.py:78-97), 128 B per problem instead of 1 712 B."""
        self.engine, self.downwash, self.depth, self.longlist = engine, downwash, int(depth), longlist
        self.lib = engine.lib
        self._p = C.c_void_p()
        mlp = downwash._h if downwash is not None else None
        import torch

        with torch.cuda.device(engine.device):
            if longlist is None:
                _lib.check(self.lib.ndp_pipeline_create(engine._h, mlp, float(r_horiz), self.depth, C.byref(self._p)), "ndp_pipeline_create")
            else:
                assert longlist.with_other == (downwash is not None)
                _lib.check(self.lib.ndp_pipeline_create_ll(engine._h, mlp, float(r_horiz), self.depth, longlist._l, C.byref(self._p)),
                           "ndp_pipeline_create_ll")
        B, N = engine.batch, engine.N
        nodes, nodes_u = (1, 1) if longlist is not None else (N + 1, N)
        np_dt = np.float32 if engine.cfg.precision == _lib.NDP_F32 else np.float64
        self.slots = []
        for s in range(self.depth):
            ptrs = [C.c_void_p() for _ in range(7)]
            _lib.check(self.lib.ndp_pipeline_buffers(self._p, s, *[C.byref(p) for p in ptrs]), "ndp_pipeline_buffers")

            def view(p, shape, dt=np_dt):
                if not p.value:
                    return None
                n = int(np.prod(shape))
                buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(p.value)
                return np.frombuffer(buf, dtype=dt).reshape(shape)

            sl = StepSlot()
            sl.x0 = view(ptrs[0], (B, NX))
            sl.xr = view(ptrs[1], (B, nodes, NX))
            sl.ur = view(ptrs[2], (B, nodes_u, NU))
            sl.other = view(ptrs[3], (B, nodes, 6))
            sl.gate_xy = view(ptrs[4], (B, 2))
            sl.u0 = view(ptrs[5], (B, NU))
            sl.status = view(ptrs[6], (B,), np.int32)
            self.slots.append(sl)
        h2d, d2h = C.c_int64(), C.c_int64()
        _lib.check(self.lib.ndp_pipeline_bytes(self._p, C.byref(h2d), C.byref(d2h)), "ndp_pipeline_bytes")
        self.h2d_bytes_per_step, self.d2h_bytes_per_step = int(h2d.value), int(d2h.value)

    def submit(self, slot: int) -> None:
        _lib.check(self.lib.ndp_pipeline_submit(self._p, slot), "ndp_pipeline_submit")

    def wait(self, slot: int) -> StepSlot:
        _lib.check(self.lib.ndp_pipeline_wait(self._p, slot), "ndp_pipeline_wait")
        return self.slots[slot]

    def step(self, slot: int = 0) -> StepSlot:
        """Synchronous step (latency path): submit + wait."""
        self.submit(slot)
        return self.wait(slot)

    def close(self):
        if getattr(self, "_p", None) is not None and self._p:
            self.lib.ndp_pipeline_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
