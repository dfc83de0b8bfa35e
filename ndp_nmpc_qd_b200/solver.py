"""Host side of the batched NMPC engine: the acados-style solver surface.

`BatchedOcpSolver` mirrors the part of acados_template.AcadosOcpSolver the reference uses
(`.N`, `.set`, `.get`, `.solve_for_x0`, `.status`; nmpc_body_rate_ctl.py:84-112,
nmpc_node.py:119,235-237) on top of the C ABI in include/ndp_nmpc.h, and adds batched tensor
variants.  PyTorch supplies device memory, streams and pinned host buffers only; all arithmetic
runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional

import numpy as np
import torch

from . import _lib
from .params import nmpc_params as CP

NX, NU, NY, NPS = 10, 4, 14, 8


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(stream=None, device=None):
    # the current stream OF THE OBJECT'S DEVICE (not of whatever device is current in the calling thread)
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return C.c_void_p(s.cuda_stream)


class Engine:
    """Thin, zero-copy wrapper of the C ABI: every tensor argument is a CUDA tensor in the
    engine precision; nothing is synchronised."""

    def __init__(self, batch: int = 1, N: int = CP.N_node, T: Optional[float] = None, np_: int = 4,
                 precision: str = "f32", device: str | torch.device = "cuda:0", **overrides):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.NdpError("CUDA device required: the NMPC engine has no CPU fallback")
        self.device = torch.device(device)
        self.dtype = torch.float32 if precision == "f32" else torch.float64
        cfg = _lib.NdpConfig()
        self.lib.ndp_default_config(C.byref(cfg))
        cfg.N, cfg.batch, cfg.np = int(N), int(batch), int(np_)
        cfg.precision = _lib.NDP_F32 if precision == "f32" else _lib.NDP_F64
        # th_pred stays 0.1 s when the horizon is lengthened
        cfg.T = float(T) if T is not None else CP.th_pred * N
        for k, v in overrides.items():
            cur = getattr(cfg, k)
            if hasattr(cur, "__len__"):
                for i, x in enumerate(v):
                    cur[i] = float(x)
            else:
                setattr(cfg, k, type(cur)(v))
        self.cfg = cfg
        self.N, self.batch, self.np = int(N), int(batch), int(np_)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ndp_create(C.byref(cfg), C.byref(self._h)), "ndp_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.ndp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, t: torch.Tensor, shape):
        assert t.is_cuda and t.dtype == self.dtype and t.is_contiguous(), "CUDA tensor in engine precision required"
        assert tuple(t.shape) == tuple(shape), f"expected shape {tuple(shape)}, got {tuple(t.shape)}"

    def n_stages(self, field: str) -> int:
        return self.N if field == "u" else self.N + 1

    def dim(self, field: str, stage: int) -> int:
        if field == "x":
            return NX
        if field == "u":
            return NU
        if field == "yref":
            return NX if stage == self.N else NY
        return self.np

    def set_stage(self, stage: int, field: str, value: torch.Tensor, stream=None):
        self._chk(value, (self.batch, self.dim(field, stage)))
        _lib.check(self.lib.ndp_set(self._h, _lib.FIELDS[field], stage, _ptr(value), value.shape[1], _stream_ptr(stream, self.device)), "ndp_set")

    def get_stage(self, stage: int, field: str, out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.batch, self.dim(field, stage)), dtype=self.dtype, device=self.device)
        self._chk(out, (self.batch, self.dim(field, stage)))
        _lib.check(self.lib.ndp_get(self._h, _lib.FIELDS[field], stage, _ptr(out), out.shape[1], _stream_ptr(stream, self.device)), "ndp_get")
        return out

    def _flat_len(self, field: str) -> int:
        ns = self.n_stages(field)
        return (ns - 1) * self.dim(field, 0) + self.dim(field, ns - 1)

    def set_all(self, field: str, value: torch.Tensor, stream=None):
        """value: [B, n_stages, dim] ('yref': flat [B, N*14+10])."""
        assert value.is_cuda and value.dtype == self.dtype and value.is_contiguous()
        assert value.numel() == self.batch * self._flat_len(field)
        _lib.check(self.lib.ndp_set(self._h, _lib.FIELDS[field], -1, _ptr(value), 0, _stream_ptr(stream, self.device)), "ndp_set")

    def get_all(self, field: str, stream=None) -> torch.Tensor:
        ns, d = self.n_stages(field), self.dim(field, 0)
        if field == "yref":
            out = torch.empty((self.batch, self._flat_len(field)), dtype=self.dtype, device=self.device)
        else:
            out = torch.empty((self.batch, ns, d), dtype=self.dtype, device=self.device)
        _lib.check(self.lib.ndp_get(self._h, _lib.FIELDS[field], -1, _ptr(out), 0, _stream_ptr(stream, self.device)), "ndp_get")
        return out

    def reset(self, xr: torch.Tensor, ur: torch.Tensor, stream=None):
        self._chk(xr, (self.batch, self.N + 1, NX))
        self._chk(ur, (self.batch, self.N, NU))
        _lib.check(self.lib.ndp_reset(self._h, _ptr(xr), _ptr(ur), _stream_ptr(stream, self.device)), "ndp_reset")

    def set_reference(self, xr: torch.Tensor, ur: torch.Tensor, f: Optional[torch.Tensor] = None, stream=None):
        self._chk(xr, (self.batch, self.N + 1, NX))
        self._chk(ur, (self.batch, self.N, NU))
        if f is not None:
            self._chk(f, (self.batch, self.N + 1, 3))
        _lib.check(self.lib.ndp_set_reference(self._h, _ptr(xr), _ptr(ur), _ptr(f), _stream_ptr(stream, self.device)), "ndp_set_reference")

    def solve(self, x0: torch.Tensor, u0: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        self._chk(x0, (self.batch, NX))
        if u0 is None:
            u0 = torch.empty((self.batch, NU), dtype=self.dtype, device=self.device)
        self._chk(u0, (self.batch, NU))
        _lib.check(self.lib.ndp_solve(self._h, _ptr(x0), _ptr(u0), _stream_ptr(stream, self.device)), "ndp_solve")
        return u0

    def update(self, x0: torch.Tensor, xr: torch.Tensor, ur: torch.Tensor, f: Optional[torch.Tensor] = None,
               u0: Optional[torch.Tensor] = None, stream=None, f_from_prev_kernel: bool = False) -> torch.Tensor:
        """controller.update() in one launch: reference upload fused into the RTI solve.
        f_from_prev_kernel: f is written by the kernel launched just before on this stream (DownwashNN) and all other
        inputs are older -- the solve is then a programmatic dependent launch that stages its inputs while that kernel
        drains (NDP_UPDATE_F_FROM_PREVIOUS_KERNEL, include/ndp_nmpc.h)."""
        self._chk(x0, (self.batch, NX))
        self._chk(xr, (self.batch, self.N + 1, NX))
        self._chk(ur, (self.batch, self.N, NU))
        if f is not None:
            self._chk(f, (self.batch, self.N + 1, 3))
        if u0 is None:
            u0 = torch.empty((self.batch, NU), dtype=self.dtype, device=self.device)
        self._chk(u0, (self.batch, NU))
        _lib.check(self.lib.ndp_update_ex(self._h, _ptr(x0), _ptr(xr), _ptr(ur), _ptr(f), _ptr(u0), 1 if (f_from_prev_kernel and f is not None) else 0,
                                          _stream_ptr(stream, self.device)), "ndp_update")
        return u0

    def status(self, out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.batch,), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.ndp_status(self._h, _ptr(out), _stream_ptr(stream, self.device)), "ndp_status")
        return out

    def stats(self, stream=None) -> torch.Tensor:
        """int32 [B, 4]: Riccati factorisations, IPM iterations, active-set rounds, active bounds."""
        out = torch.empty((self.batch, 4), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.ndp_stats(self._h, _ptr(out), _stream_ptr(stream, self.device)), "ndp_stats")
        return out

    @property
    def launch_count(self) -> int:
        return int(self.lib.ndp_launch_count(self._h))

    def kernel_timing(self, enable: bool = True) -> None:
        """Measurement aid (ndp_kernel_timing): events around the two kernels of every solve."""
        _lib.check(self.lib.ndp_kernel_timing(self._h, 1 if enable else 0), "ndp_kernel_timing")

    def last_kernel_ms(self):
        """(nominal kernel ms, constrained kernel ms) of the last solve; waits for it."""
        a, b = C.c_float(), C.c_float()
        _lib.check(self.lib.ndp_last_kernel_ms(self._h, C.byref(a), C.byref(b)), "ndp_last_kernel_ms")
        return float(a.value), float(b.value)


class BatchedOcpSolver:
    """acados-style surface over `Engine` with host mirrors, so that the per-stage
    `set(stage, field, value)` calls of the reference controller cost a numpy copy each and the
    whole record crosses PCIe once per solve.

    batch == 1 reproduces AcadosOcpSolver's shapes exactly (`set` takes [dim], `get` returns a
    fresh [dim] float64 array, `solve_for_x0` returns [4], `status` is an int); batch > 1 adds a
    leading batch axis everywhere.
    """

    def __init__(self, batch: int = 1, N: int = CP.N_node, T: Optional[float] = None, np_: int = 4,
                 precision: str = "f32", device: str | torch.device = "cuda:0", **overrides):
        self.engine = Engine(batch, N, T, np_, precision, device, **overrides)
        self.batch, self._N, self.np = batch, N, np_
        self.device, self.dtype = self.engine.device, self.engine.dtype
        self._lock = threading.RLock()  # get() from a viz thread while solve() runs (nmpc_node.py:233-237)
        B = batch
        NPS = 8  # device stride of a parameter record: q_r(4) f(3) pad
        # host mirrors in pinned memory, in the layouts the library stores (ndp_solve_host copies them as they are):
        # yref [B, N+1, 14] (terminal row: first 10 used), p [B, N+1, 8], x0 [B, 10]; outputs u0 [B, 4], status [B]
        self._pin_x0 = torch.zeros((B, NX), dtype=self.dtype).pin_memory()
        self._pin_yref = torch.zeros((B, N + 1, NY), dtype=self.dtype).pin_memory()
        self._pin_p = torch.zeros((B, N + 1, NPS), dtype=self.dtype).pin_memory()
        self._pin_u0 = torch.zeros((B, NU), dtype=self.dtype).pin_memory()
        self._pin_status = torch.zeros((B,), dtype=torch.int32).pin_memory()
        self._h_x0, self._h_yref, self._h_pfull = self._pin_x0.numpy(), self._pin_yref.numpy(), self._pin_p.numpy()
        self._h_p = self._h_pfull[:, :, :np_]
        self._ho_u0 = self._pin_u0.numpy()
        # iterate mirror (uploaded after reset / set of x, u; read back on demand)
        self._sz_X, self._sz_U = B * (N + 1) * NX, B * N * NU
        self._pin_it = torch.zeros(self._sz_X + self._sz_U, dtype=self.dtype).pin_memory()
        self._pin_out = torch.zeros(self._sz_X + self._sz_U, dtype=self.dtype).pin_memory()
        self._d_it = torch.zeros_like(self._pin_it, device=self.device)
        self._d_out = torch.zeros_like(self._pin_out, device=self.device)

        def regions(t, sizes, shapes):
            out, o = [], 0
            for n, shp in zip(sizes, shapes):
                out.append(t[o:o + n].view(*shp))
                o += n
            return out

        it_shapes = [(B, N + 1, NX), (B, N, NU)]
        self._h_X, self._h_U = (r.numpy() for r in regions(self._pin_it, (self._sz_X, self._sz_U), it_shapes))
        self._dv_X, self._dv_U = regions(self._d_it, (self._sz_X, self._sz_U), it_shapes)
        self._ho_X, self._ho_U = (r.numpy() for r in regions(self._pin_out, (self._sz_X, self._sz_U), it_shapes))
        self._do_X, self._do_U = regions(self._d_out, (self._sz_X, self._sz_U), it_shapes)
        self._dirty_ref = True
        self._dirty_it = False
        self._it_stale = False  # the device iterate is newer than the host mirror (fetched on demand by get / set)
        self._status = np.zeros(B, np.int32)
        self.stream = torch.cuda.Stream(device=self.device)
        self._hp = [C.c_void_p(t.data_ptr()) for t in (self._pin_x0, self._pin_yref, self._pin_p, self._pin_u0, self._pin_status)]

    def _ydim(self, stage: int) -> int:
        return NX if stage == self._N else NY

    # ---- acados surface ----
    @property
    def N(self) -> int:
        return self._N

    def _sync_iterate(self) -> None:
        """Bring the host mirror of (X, U) up to date with the device iterate.  The per-step path only reads u0 and
        the status back; the predicted trajectory crosses PCIe when somebody asks for it (the node's viz timer,
        nmpc_node.py:233-237, or a per-stage set of x / u)."""
        if not self._it_stale:
            return
        e = self.engine
        with torch.cuda.stream(self.stream):
            _lib.check(e.lib.ndp_get(e._h, _lib.FIELD_X, -1, _ptr(self._do_X), 0, _stream_ptr(self.stream)), "ndp_get")
            _lib.check(e.lib.ndp_get(e._h, _lib.FIELD_U, -1, _ptr(self._do_U), 0, _stream_ptr(self.stream)), "ndp_get")
            self._pin_out.copy_(self._d_out, non_blocking=True)
        self.stream.synchronize()
        self._h_X[...] = self._ho_X
        self._h_U[...] = self._ho_U
        self._it_stale = False

    def set(self, stage: int, field: str, value) -> None:
        """solver.set(stage, field, value) -- nmpc_body_rate_ctl.py:89-91,97-104."""
        v = np.asarray(value)
        with self._lock:
            if field in ("x", "u"):
                self._sync_iterate()
            if field == "x":
                self._h_X[:, stage, :] = v.reshape(-1, NX)
                self._dirty_it = True
            elif field == "u":
                self._h_U[:, stage, :] = v.reshape(-1, NU)
                self._dirty_it = True
            elif field == "yref":
                d = self._ydim(stage)
                self._h_yref[:, stage, :d] = v.reshape(-1, d)
                self._dirty_ref = True
            elif field == "p":
                self._h_p[:, stage, :] = v.reshape(-1, self.np)
                self._dirty_ref = True
            else:
                raise Exception(f"AcadosOcpSolver.set(): {field} is not a valid argument.")

    def get(self, stage: int, field: str) -> np.ndarray:
        """solver.get(stage, field): a fresh float64 copy (the node mutates it, nmpc_node.py:237-238)."""
        with self._lock:
            if field in ("x", "u"):
                self._sync_iterate()
            if field == "x":
                out = self._h_X[:, stage, :]
            elif field == "u":
                out = self._h_U[:, stage, :]
            elif field == "yref":
                out = self._h_yref[:, stage, :self._ydim(stage)]
            elif field == "p":
                out = self._h_p[:, stage, :]
            else:
                raise Exception(f"AcadosOcpSolver.get(): {field} is not a valid argument.")
            out = np.array(out, dtype=np.float64)
        return out[0] if self.batch == 1 else out

    def solve_for_x0(self, x0_bar) -> np.ndarray:
        """u0 = solver.solve_for_x0(x0) -- nmpc_body_rate_ctl.py:107.  One C call (ndp_solve_host): reference upload if
        it changed, one SQP_RTI step, u0 and status back; the predicted trajectory is read back on demand."""
        with self._lock:
            e = self.engine
            self._h_x0[:, :] = np.asarray(x0_bar).reshape(-1, NX)
            if self._dirty_it:
                with torch.cuda.stream(self.stream):
                    self._d_it.copy_(self._pin_it, non_blocking=True)
                    e.set_all("x", self._dv_X, stream=self.stream)
                    e.set_all("u", self._dv_U, stream=self.stream)
                self.stream.synchronize()
                self._dirty_it = False
            with torch.cuda.device(self.device):
                _lib.check(e.lib.ndp_solve_host(e._h, self._hp[0], self._hp[1], self._hp[2], 1 if self._dirty_ref else 0,
                                                self._hp[3], self._hp[4]), "ndp_solve_host")
            self._dirty_ref = False
            self._it_stale = True
            self._status = self._pin_status.numpy().copy()
            u0 = np.array(self._ho_u0, dtype=np.float64)
        return u0[0] if self.batch == 1 else u0

    @property
    def status(self):
        """solver.status -- nmpc_body_rate_ctl.py:109 (int for batch 1, int32 array otherwise)."""
        return int(self._status[0]) if self.batch == 1 else self._status

    def get_stats(self, _field: str = "stats"):
        return self.engine.stats().cpu().numpy()

    # ---- batched extensions ----
    def set_all(self, field: str, value) -> None:
        v = np.asarray(value)
        with self._lock:
            self._sync_iterate()
            if field == "x":
                self._h_X[...] = v.reshape(self.batch, self._N + 1, NX)
                self._dirty_it = True
            elif field == "u":
                self._h_U[...] = v.reshape(self.batch, self._N, NU)
                self._dirty_it = True
            else:
                raise ValueError("set_all supports 'x' and 'u'; use set_reference for yref / p")

    def get_all(self, field: str) -> np.ndarray:
        with self._lock:
            self._sync_iterate()
            src = {"x": self._h_X, "u": self._h_U}[field]
            return np.array(src, dtype=np.float64)

    def set_reference(self, xr, ur, f=None) -> None:
        """The 42 set() calls of controller.update() at once (nmpc_body_rate_ctl.py:95-104)."""
        xr = np.asarray(xr).reshape(self.batch, self._N + 1, NX)
        ur = np.asarray(ur).reshape(self.batch, self._N, NU)
        with self._lock:
            N = self._N
            self._h_yref[:, :N, :NX] = xr[:, :N, :]
            self._h_yref[:, :N, NX:] = ur
            self._h_yref[:, N, :NX] = xr[:, N, :]
            self._h_p[:, :, :4] = xr[:, :, 6:10]
            if f is not None:
                if self.np != 7:
                    raise ValueError("disturbance forces need np = 7 (NDPNMPCBodyRateController)")
                self._h_p[:, :, 4:7] = np.asarray(f).reshape(self.batch, self._N + 1, 3)
            self._dirty_ref = True

    def reset(self, xr, ur) -> None:
        self.set_all("x", xr)
        self.set_all("u", ur)
