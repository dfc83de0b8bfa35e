"""RefGen: batched NMPC reference horizons on the device (ndp_refgen_* in include/ndp_nmpc.h).

Device-side counterpart of NMPCRefPublisher.get_nmpc_pts (ndp_nmpc/scripts/pt_pub/pt_publisher.py:78-103):
for every problem b the horizon xr[b, k], ur[b, k] is the planner's polynomial trajectory pushed through
the differential-flatness map at t0[b] + k * th_pred.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _lib
from ..params import nmpc_params as CP
from .min_snap import Trajectory


def _sp(stream=None, device=None):
    # the current stream OF THE OBJECT'S DEVICE (not of whatever device is current in the calling thread)
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return C.c_void_p(s.cuda_stream)


class RefGen:
    def __init__(self, trajectories: Sequence[Trajectory], device="cuda:0"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.NdpError("CUDA device required: RefGen has no CPU fallback")
        self.device = torch.device(device)
        self.trajectories = list(trajectories)
        seg_off = np.zeros(len(self.trajectories) + 1, dtype=np.int32)
        for i, tr in enumerate(self.trajectories):
            seg_off[i + 1] = seg_off[i] + len(tr.t_cum) - 1
        cat = lambda f: np.ascontiguousarray(np.concatenate([np.asarray(f(tr), dtype=np.float64).reshape(-1) for tr in self.trajectories]))
        arrs = [cat(lambda tr: tr.t_cum), cat(lambda tr: tr.cx), cat(lambda tr: tr.cy), cat(lambda tr: tr.cz), cat(lambda tr: tr.cyaw),
                cat(lambda tr: tr.final_pt)]
        dp = C.POINTER(C.c_double)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ndp_refgen_create(len(self.trajectories), seg_off.ctypes.data_as(C.POINTER(C.c_int32)),
                                                  *[a.ctypes.data_as(dp) for a in arrs], C.byref(self._h)), "ndp_refgen_create")

    def horizon(self, t0: torch.Tensor, traj_id: Optional[torch.Tensor] = None, N: int = CP.N_node, th_pred: float = CP.th_pred,
                offset: Optional[torch.Tensor] = None, dtype=torch.float32, xr: Optional[torch.Tensor] = None,
                ur: Optional[torch.Tensor] = None, stream=None):
        """t0 [B] float64 CUDA (trajectory time of node 0), traj_id [B] int32 or None, offset [B,3] float64 or None.
        Returns xr [B,N+1,10], ur [B,N,4] in `dtype`."""
        assert t0.is_cuda and t0.dtype == torch.float64 and t0.is_contiguous()
        B = t0.numel()
        if traj_id is not None:
            assert traj_id.is_cuda and traj_id.dtype == torch.int32 and traj_id.numel() == B and traj_id.is_contiguous()
        if offset is not None:
            assert offset.is_cuda and offset.dtype == torch.float64 and offset.is_contiguous() and offset.numel() == 3 * B
        if xr is None:
            xr = torch.empty((B, N + 1, 10), dtype=dtype, device=t0.device)
        if ur is None:
            ur = torch.empty((B, N, 4), dtype=dtype, device=t0.device)
        assert xr.dtype == ur.dtype and xr.is_contiguous() and ur.is_contiguous() and xr.numel() == B * (N + 1) * 10 and ur.numel() == B * N * 4
        prec = _lib.NDP_F32 if xr.dtype == torch.float32 else _lib.NDP_F64
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        _lib.check(self.lib.ndp_refgen_horizon(self._h, prec, B, p(traj_id), p(t0), int(N), float(th_pred), p(offset), p(xr), p(ur), _sp(stream, self.device)),
                   "ndp_refgen_horizon")
        return xr, ur

    @property
    def launch_count(self) -> int:
        return int(self.lib.ndp_refgen_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ndp_refgen_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
