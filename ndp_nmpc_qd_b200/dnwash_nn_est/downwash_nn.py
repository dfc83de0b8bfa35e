"""DownwashNN: the 6-128-64-128-3 ReLU MLP of the reference, on hand-written CUDA kernels.

Reference: dnwash_nn_est/downwash_nn.py:10-29 (`update(other_pred_x, ego_pred_x)` ->
float32 [21,3]; features = (other - ego)[:, 0:6] in float64, cast to float32), net
nn_net.py:7-18, deployed weights downwash_nn.py:15.  Batched variants fuse the feature
construction, the 1 m horizontal gate of ndp_nmpc_leader_node.py:65-76 and, for swarms, the sum
over gated neighbours.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..params import downwash_params as DP

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_WEIGHTS = os.path.join(_HERE, "nn_model", "128-64-128_WBias_SN=4_epoch=20000_test_loss=1.0221.npz")
_KEYS = ("0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias", "6.weight", "6.bias")
_ALT = {"fc1": "0", "fc2": "2", "fc3": "4", "fc4": "6"}  # older checkpoints of the reference


def load_weights(path: str) -> dict:
    """Read a state_dict from the reference's torch pickle (.pkl/.pt) or from the .npz export."""
    if path.endswith(".npz"):
        raw = dict(np.load(path))
    else:
        raw = {k: v.detach().cpu().numpy() for k, v in torch.load(path, map_location="cpu", weights_only=True).items()}
    out = {}
    for k, v in raw.items():
        head, tail = k.split(".")
        out[f"{_ALT.get(head, head)}.{tail}"] = np.ascontiguousarray(v, dtype=np.float32)
    shapes = [(128, 6), (128,), (64, 128), (64,), (128, 64), (128,), (3, 128), (3,)]
    for k, s in zip(_KEYS, shapes):
        if k not in out or out[k].shape != s:
            raise ValueError(f"{path}: missing or mis-shaped parameter {k} (want {s})")
    return out


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _sp(stream=None, device=None):
    # the current stream OF THE OBJECT'S DEVICE (not of whatever device is current in the calling thread)
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return C.c_void_p(s.cuda_stream)


class DownwashNN:
    PATH_AUTO, PATH_FP32, PATH_TENSOR, PATH_ROWS = 0, 1, 2, 3

    def __init__(self, weights: Optional[str] = None, device: str | torch.device = "cuda:0"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.NdpError("CUDA device required: DownwashNN has no CPU fallback")
        self.device = torch.device(device)
        # the reference loads a cwd-relative path after the controller's chdir (downwash_nn.py:15);
        # resolve absolutely but accept that spelling too
        path = weights or DEFAULT_WEIGHTS
        if not os.path.isabs(path) and not os.path.exists(path):
            path = os.path.join(_HERE, os.path.basename(os.path.dirname(path)), os.path.basename(path))
        self.weights = load_weights(path)
        fp = C.POINTER(C.c_float)
        args = [self.weights[k].ctypes.data_as(fp) for k in _KEYS]
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ndp_mlp_create(*args, C.byref(self._h)), "ndp_mlp_create")
        self._pin_in = torch.zeros((2, 21, 10), dtype=torch.float64).pin_memory()
        self._pin_out = torch.zeros((21, 3), dtype=torch.float64).pin_memory()
        self.stream = torch.cuda.Stream(device=self.device)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ndp_mlp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference interface ----
    def update(self, other_pred_x: np.ndarray, ego_pred_x: np.ndarray) -> np.ndarray:
        """float32 [n,3] forces in Newton for one neighbour (downwash_nn.py:21-29)."""
        n = other_pred_x.shape[0]
        if n != self._pin_in.shape[1]:
            self._pin_in = torch.zeros((2, n, 10), dtype=torch.float64).pin_memory()
            self._pin_out = torch.zeros((n, 3), dtype=torch.float64).pin_memory()
        h = self._pin_in.numpy()
        h[0] = ego_pred_x
        h[1] = other_pred_x
        # latency path: the row kernel reads the pinned host arrays and writes the forces over PCIe itself (unified
        # addressing makes pinned host memory device-accessible at the same address) -- one launch, no copy operations
        sp = C.c_void_p(self.stream.cuda_stream)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ndp_mlp_forward_pairs(self._h, _lib.NDP_F64, 1, n, C.c_void_p(self._pin_in[0].data_ptr()),
                                                      C.c_void_p(self._pin_in[1].data_ptr()), None, float(DP.r_horiz),
                                                      C.c_void_p(self._pin_out.data_ptr()), 0, 0, sp), "ndp_mlp_forward_pairs")
        self.stream.synchronize()
        return self._pin_out.numpy().astype(np.float32)

    # ---- batched device API ----
    def forward_rows(self, x: torch.Tensor, path: int = 0, stream=None) -> torch.Tensor:
        """The nn.Sequential itself: x [M,6] float32 CUDA -> [M,3] float32."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[-1] == 6
        out = torch.empty((x.shape[0], 3), dtype=torch.float32, device=x.device)
        _lib.check(self.lib.ndp_mlp_forward_rows(self._h, x.shape[0], _ptr(x), _ptr(out), path, _sp(stream, self.device)), "ndp_mlp_forward_rows")
        return out

    def forward_pairs(self, ego: torch.Tensor, other: torch.Tensor, gate_xy: Optional[torch.Tensor] = None,
                      r_horiz: float = DP.r_horiz, out: Optional[torch.Tensor] = None, accumulate: bool = False,
                      path: int = 0, stream=None) -> torch.Tensor:
        """ego/other [P,n,10] (float32 or float64 CUDA); gate_xy [P,2] ego odometry position or None.
        Returns f [P,n,3] in the same dtype."""
        assert ego.is_cuda and ego.shape == other.shape and ego.dtype == other.dtype
        assert ego.is_contiguous() and other.is_contiguous() and ego.shape[-1] == 10
        P, n = ego.shape[0], ego.shape[1]
        prec = _lib.NDP_F32 if ego.dtype == torch.float32 else _lib.NDP_F64
        if out is None:
            out = torch.empty((P, n, 3), dtype=ego.dtype, device=ego.device)
        assert out.dtype == ego.dtype and out.is_contiguous() and out.numel() == P * n * 3
        if gate_xy is not None:
            assert gate_xy.dtype == ego.dtype and gate_xy.is_contiguous() and gate_xy.shape == (P, 2)
        _lib.check(self.lib.ndp_mlp_forward_pairs(self._h, prec, P, n, _ptr(ego), _ptr(other), _ptr(gate_xy), float(r_horiz),
                                                  _ptr(out), int(accumulate), path, _sp(stream, self.device)), "ndp_mlp_forward_pairs")
        return out

    def forward_swarm(self, traj: torch.Tensor, ego_begin: int, n_ego: int, odom_xy: Optional[torch.Tensor] = None,
                      r_horiz: float = DP.r_horiz, out_dtype=torch.float32, path: int = 0, stream=None) -> torch.Tensor:
        """traj [n_all,n,6] float32 (every quad's reference positions+velocities, e.g. all-gathered);
        returns f [n_ego,n,3] = sum over gated neighbours j != i of MLP(traj_j - traj_i)."""
        assert traj.is_cuda and traj.dtype == torch.float32 and traj.is_contiguous() and traj.shape[-1] == 6
        n_all, n = traj.shape[0], traj.shape[1]
        out = torch.empty((n_ego, n, 3), dtype=out_dtype, device=traj.device)
        prec = _lib.NDP_F32 if out_dtype == torch.float32 else _lib.NDP_F64
        if odom_xy is not None:
            assert odom_xy.dtype == torch.float32 and odom_xy.is_contiguous() and odom_xy.shape == (n_ego, 2)
        _lib.check(self.lib.ndp_mlp_forward_swarm(self._h, prec, n_all, ego_begin, n_ego, n, _ptr(traj), _ptr(odom_xy),
                                                  float(r_horiz), _ptr(out), path, _sp(stream, self.device)), "ndp_mlp_forward_swarm")
        return out

    def set_group(self, group: int) -> None:
        """forward_swarm: quads interact inside contiguous blocks of `group` only (0: all pairs)."""
        _lib.check(self.lib.ndp_mlp_set_group(self._h, int(group)), "ndp_mlp_set_group")

    @property
    def launch_count(self) -> int:
        return int(self.lib.ndp_mlp_launch_count(self._h))
