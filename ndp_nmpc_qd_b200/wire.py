"""Batched device-side forms of the glue the reference's nodes run around the controller every tick
(SURVEY.md 8f-3, 8a row a9): the PredXU topic payload (ndp_nmpc/msg/PredXU.msg; do_pub_ref nmpc_node.py:116-133,
consumers nmpc_follower_node.py:57-74 and ndp_nmpc_leader_node.py:60-76) and the hover-throttle estimator hook
(nmpc_node.py:251-253).  Thin wrappers of the C ABI (include/ndp_nmpc.h); torch only owns the buffers."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .params import estimator_params as EP


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _sp(stream, device):
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return C.c_void_p(s.cuda_stream)


def _prec(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.NDP_F32
    if t.dtype == torch.float64:
        return _lib.NDP_F64
    raise TypeError("float32 or float64 tensor required")


def predxu_len(N: int) -> int:
    """float64 elements of one quadrotor's PredXU payload: (N+1) x-rows of 10, then N u-rows of 4."""
    return int(_lib.load().ndp_predxu_len(int(N)))


def predxu_pack(xr: torch.Tensor, ur: torch.Tensor, out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
    """do_pub_ref (nmpc_node.py:116-133) for B quadrotors: xr [B,N+1,10], ur [B,N,4] -> float64 [B, predxu_len(N)]."""
    lib = _lib.load()
    B, N = xr.shape[0], ur.shape[1]
    assert xr.is_cuda and ur.is_cuda and xr.is_contiguous() and ur.is_contiguous() and xr.dtype == ur.dtype
    assert tuple(xr.shape) == (B, N + 1, 10) and tuple(ur.shape) == (B, N, 4)
    if out is None:
        out = torch.empty((B, predxu_len(N)), dtype=torch.float64, device=xr.device)
    assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.numel() == B * predxu_len(N)
    _lib.check(lib.ndp_predxu_pack(_prec(xr), B, N, _p(xr), _p(ur), _p(out), _sp(stream, xr.device)), "ndp_predxu_pack")
    return out


def predxu_unpack(msg: torch.Tensor, N: int, dtype=torch.float32, offset: Optional[torch.Tensor] = None,
                  xr: Optional[torch.Tensor] = None, ur: Optional[torch.Tensor] = None, stream=None):
    """FollowerNode.sub_pred_callback (nmpc_follower_node.py:57-74) for B quadrotors: float64 [B, predxu_len(N)] ->
    (xr [B,N+1,10], ur [B,N,4]) in `dtype`, with the formation offset [B,3] (float64) added to every x row's position."""
    lib = _lib.load()
    B = msg.shape[0]
    assert msg.is_cuda and msg.dtype == torch.float64 and msg.is_contiguous() and msg.numel() == B * predxu_len(N)
    if xr is None:
        xr = torch.empty((B, N + 1, 10), dtype=dtype, device=msg.device)
    if ur is None:
        ur = torch.empty((B, N, 4), dtype=dtype, device=msg.device)
    if offset is not None:
        assert offset.is_cuda and offset.dtype == torch.float64 and offset.is_contiguous() and offset.numel() == B * 3
    _lib.check(lib.ndp_predxu_unpack(_prec(xr), B, N, _p(msg), _p(offset), _p(xr), _p(ur), _sp(stream, msg.device)), "ndp_predxu_unpack")
    return xr, ur


class BatchedHoverThrottleEstimator:
    """HoverThrottleEstimator (hv_throttle_est/hover_throttle_estimator.py:15-53) for n quadrotors, state on the
    device.  `k_throttle` [n] float64 is what nmpc_u_2_att_tgt divides by (MulQuadrotors.cmd_from_u0_dev)."""

    def __init__(self, n: int, ts: float = EP.ts_est, device="cuda:0"):
        self.lib = _lib.load()
        self.n, self.ts, self.device = int(n), float(ts), torch.device(device)
        self.est = torch.empty((self.n, 8), dtype=torch.float64, device=self.device)
        self.k_throttle = torch.empty((self.n,), dtype=torch.float64, device=self.device)
        self.reset()

    def reset(self, stream=None):
        _lib.check(self.lib.ndp_hover_throttle_init(self.n, _p(self.est), _p(self.k_throttle), _sp(stream, self.device)), "ndp_hover_throttle_init")

    def update(self, vz: torch.Tensor, throttle: torch.Tensor, stream=None) -> torch.Tensor:
        """vz / throttle: float64 CUDA views with one element per quadrotor (any element stride, e.g.
        plant_state[:, 15, 0] and cmd[:, 3, 0]).  Returns k_throttle [n]."""
        for t in (vz, throttle):
            assert t.is_cuda and t.dtype == torch.float64 and t.dim() == 1 and t.shape[0] == self.n
        _lib.check(self.lib.ndp_hover_throttle_update(self.n, self.ts, _p(vz), int(vz.stride(0)) if self.n > 1 else 1, _p(throttle),
                                                      int(throttle.stride(0)) if self.n > 1 else 1, _p(self.est), _p(self.k_throttle),
                                                      _sp(stream, self.device)), "ndp_hover_throttle_update")
        return self.k_throttle
