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


class StepSlot:
    """numpy views of one slot's pinned host arrays."""
    __slots__ = ("x0", "xr", "ur", "other", "gate_xy", "u0", "status")


class HostStepPipeline:
    def __init__(self, engine: Engine, downwash=None, depth: int = 2, r_horiz: float = DP.r_horiz):
        """engine: `Engine` (np_=7 when `downwash` is given); downwash: `DownwashNN` or None."""
        self.engine, self.downwash, self.depth = engine, downwash, int(depth)
        self.lib = engine.lib
        self._p = C.c_void_p()
        mlp = downwash._h if downwash is not None else None
        import torch

        with torch.cuda.device(engine.device):
            _lib.check(self.lib.ndp_pipeline_create(engine._h, mlp, float(r_horiz), self.depth, C.byref(self._p)), "ndp_pipeline_create")
        B, N = engine.batch, engine.N
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
            sl.xr = view(ptrs[1], (B, N + 1, NX))
            sl.ur = view(ptrs[2], (B, N, NU))
            sl.other = view(ptrs[3], (B, N + 1, 6))
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
