"""MulQuadrotors: the dop_sim batched quadrotor plant on hand-written CUDA kernels (float64).

Mirror of /root/reference/dop_sim/scripts/quadrotor/mul_quadrotors.py:19-50 -- same constructor
arguments and `forward(ts_sim, ego_states[n,35,1], body_rate_cmd[n,4,1]) -> ego_states` (the state
tensor is updated in place and returned, as the reference's dynamics does, qd_dynamics.py:75-98) -- so
dop_qd_node.py:200-225 can construct and call it unchanged (no torch.jit.script needed: the module is
not an nn.Module graph but a thin wrapper over ndp_plant_* in include/ndp_nmpc.h).  There is no CPU
fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib


def _sp(stream=None, device=None):
    # the current stream OF THE OBJECT'S DEVICE (not of whatever device is current in the calling thread)
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return C.c_void_p(s.cuda_stream)


class MulQuadrotors:
    def __init__(self, num_agent: int, ts_sim: float, ts_control: float, dtype=torch.float64, has_downwash=True,
                 has_motor_model=True, has_battery=True, group: int = 0, device="cuda:0"):
        if dtype != torch.float64:
            raise ValueError("the plant kernels are float64 (the reference simulator runs in float64, dop_qd_node.py:173)")
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.NdpError("CUDA device required: the plant has no CPU fallback")
        self.device = torch.device(device)
        self.num_agent, self.ts_sim, self.ts_ctl = int(num_agent), float(ts_sim), float(ts_control)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ndp_plant_create(self.num_agent, self.ts_sim, self.ts_ctl, int(has_downwash), int(has_motor_model),
                                                 int(has_battery), int(group), C.byref(self._h)), "ndp_plant_create")

    # nn.Module look-alikes used by the reference node (dop_qd_node.py:205-212)
    def requires_grad_(self, _flag=False):
        return self

    def to(self, _device):
        return self

    def _chk(self, ego_states, cmd):
        n = self.num_agent
        assert ego_states.is_cuda and ego_states.dtype == torch.float64 and ego_states.is_contiguous() and ego_states.numel() == n * 35
        assert cmd.is_cuda and cmd.dtype == torch.float64 and cmd.is_contiguous() and cmd.numel() == n * 4

    def forward(self, ts_sim: float, ego_states: torch.Tensor, body_rate_cmd: torch.Tensor, stream=None) -> torch.Tensor:
        self._chk(ego_states, body_rate_cmd)
        _lib.check(self.lib.ndp_plant_forward(self._h, float(ts_sim), C.c_void_p(ego_states.data_ptr()),
                                              C.c_void_p(body_rate_cmd.data_ptr()), _sp(stream, self.device)), "ndp_plant_forward")
        return ego_states

    __call__ = forward

    def reset(self, stream=None):
        _lib.check(self.lib.ndp_plant_reset(self._h, _sp(stream, self.device)), "ndp_plant_reset")

    # ---- glue to the batched NMPC engine (device side, no host round trip) ----
    def nmpc_x0(self, ego_states: torch.Tensor, out: torch.Tensor, stream=None) -> torch.Tensor:
        """odom_2_nmpc_x (pt_publisher.py:106-122): x0 [n,10] = (p, v, qw, qx, qy, qz) in out's dtype."""
        prec = _lib.NDP_F32 if out.dtype == torch.float32 else _lib.NDP_F64
        assert out.is_cuda and out.is_contiguous() and out.numel() == self.num_agent * 10
        _lib.check(self.lib.ndp_plant_nmpc_x0(self.num_agent, C.c_void_p(ego_states.data_ptr()), prec, C.c_void_p(out.data_ptr()), _sp(stream, self.device)),
                   "ndp_plant_nmpc_x0")
        return out

    def cmd_from_u0(self, u0: torch.Tensor, cmd: torch.Tensor, mass: float, k_throttle: float, stream=None) -> torch.Tensor:
        """nmpc_u_2_att_tgt (nmpc_node.py:273-283): body rates = u0[:, 0:3], thrust = u0[:, 3] * mass / k_throttle."""
        prec = _lib.NDP_F32 if u0.dtype == torch.float32 else _lib.NDP_F64
        assert u0.is_cuda and u0.is_contiguous() and cmd.is_cuda and cmd.dtype == torch.float64 and cmd.is_contiguous()
        _lib.check(self.lib.ndp_plant_cmd_from_u0(self.num_agent, prec, C.c_void_p(u0.data_ptr()), float(mass), float(k_throttle),
                                                  C.c_void_p(cmd.data_ptr()), _sp(stream, self.device)), "ndp_plant_cmd_from_u0")
        return cmd

    def cmd_from_u0_dev(self, u0: torch.Tensor, cmd: torch.Tensor, mass: float, k_throttle: torch.Tensor, stream=None) -> torch.Tensor:
        """nmpc_u_2_att_tgt with one hover-throttle estimate per quadrotor (k_throttle float64 [n] on the device)."""
        prec = _lib.NDP_F32 if u0.dtype == torch.float32 else _lib.NDP_F64
        assert u0.is_cuda and u0.is_contiguous() and cmd.is_cuda and cmd.dtype == torch.float64 and cmd.is_contiguous()
        assert k_throttle.is_cuda and k_throttle.dtype == torch.float64 and k_throttle.is_contiguous() and k_throttle.numel() == self.num_agent
        _lib.check(self.lib.ndp_plant_cmd_from_u0_dev(self.num_agent, prec, C.c_void_p(u0.data_ptr()), float(mass), C.c_void_p(k_throttle.data_ptr()),
                                                      C.c_void_p(cmd.data_ptr()), _sp(stream, self.device)), "ndp_plant_cmd_from_u0_dev")
        return cmd

    @property
    def launch_count(self) -> int:
        return int(self.lib.ndp_plant_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ndp_plant_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
