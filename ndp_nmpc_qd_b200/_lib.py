"""ctypes binding of libndp_nmpc_b200.so (the C ABI declared in include/ndp_nmpc.h).

There is no CPU fallback: if the shared library is missing this raises, and every compute
entry point launches CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NDP_NMPC_LIB: an alternative build of the same library (A/B measurements of kernel variants, tools/ab_variants.py)
LIB_PATH = os.environ.get("NDP_NMPC_LIB") or os.path.join(_HERE, "_C", "libndp_nmpc_b200.so")

NDP_F32, NDP_F64 = 0, 1
FIELD_X, FIELD_U, FIELD_YREF, FIELD_P = 0, 1, 2, 3
FIELDS = {"x": FIELD_X, "u": FIELD_U, "yref": FIELD_YREF, "p": FIELD_P}

# every symbol include/ndp_nmpc.h declares
EXPORTS = [
    "ndp_default_config", "ndp_create", "ndp_destroy", "ndp_set", "ndp_get", "ndp_reset", "ndp_set_reference",
    "ndp_solve", "ndp_update", "ndp_status", "ndp_stats", "ndp_launch_count", "ndp_last_error", "ndp_rk4_sens",
    "ndp_mlp_create", "ndp_mlp_destroy", "ndp_mlp_forward_pairs", "ndp_mlp_forward_rows", "ndp_mlp_forward_swarm",
    "ndp_mlp_launch_count", "ndp_mlp_forward_pairs_ex", "ndp_mlp_forward_swarm_parts",
    "ndp_pipeline_create", "ndp_pipeline_destroy", "ndp_pipeline_buffers", "ndp_pipeline_submit", "ndp_pipeline_wait",
    "ndp_pipeline_bytes", "ndp_pipeline_stream", "ndp_mlp_set_pair_budget", "ndp_solve_host", "ndp_update_ex",
    "ndp_plant_create", "ndp_plant_destroy", "ndp_plant_reset", "ndp_plant_forward", "ndp_plant_autopilot", "ndp_plant_dynamics",
    "ndp_plant_nmpc_x0", "ndp_plant_cmd_from_u0", "ndp_plant_launch_count",
    "ndp_refgen_create", "ndp_refgen_destroy", "ndp_refgen_horizon", "ndp_refgen_launch_count",
    "ndp_predxu_len", "ndp_predxu_pack", "ndp_predxu_unpack", "ndp_hover_throttle_init", "ndp_hover_throttle_update",
    "ndp_plant_cmd_from_u0_dev",
    "ndp_longlist_create", "ndp_longlist_destroy", "ndp_longlist_reset", "ndp_longlist_push", "ndp_longlist_launch_count",
    "ndp_pipeline_create_ll", "ndp_kernel_timing", "ndp_last_kernel_ms", "ndp_mlp_set_group",
]


class NdpConfig(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("precision", C.c_int32), ("batch", C.c_int32), ("np", C.c_int32),
        ("T", C.c_double), ("mass", C.c_double), ("gravity", C.c_double),
        ("Q", C.c_double * 10), ("R", C.c_double * 4),
        ("u_min", C.c_double * 4), ("u_max", C.c_double * 4),
        ("v_min", C.c_double * 3), ("v_max", C.c_double * 3),
        ("ipm_max_iter", C.c_int32), ("polish_max", C.c_int32), ("ipm_tol_mu", C.c_double),
        ("active_set_first", C.c_int32), ("active_set_warm", C.c_int32),
    ]


class NdpError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NdpError(
            f"{LIB_PATH} is missing: build it with `python -m ndp_nmpc_qd_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.ndp_default_config.argtypes = [C.POINTER(NdpConfig)]
    lib.ndp_default_config.restype = None
    lib.ndp_create.argtypes = [C.POINTER(NdpConfig), C.POINTER(vp)]
    lib.ndp_destroy.argtypes = [vp]
    lib.ndp_set.argtypes = [vp, i32, i32, vp, i64, vp]
    lib.ndp_get.argtypes = [vp, i32, i32, vp, i64, vp]
    lib.ndp_reset.argtypes = [vp, vp, vp, vp]
    lib.ndp_set_reference.argtypes = [vp, vp, vp, vp, vp]
    lib.ndp_solve.argtypes = [vp, vp, vp, vp]
    lib.ndp_update.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.ndp_update_ex.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    lib.ndp_update_ex.restype = C.c_int
    lib.ndp_solve_host.argtypes = [vp, vp, vp, vp, i32, vp, vp]
    lib.ndp_solve_host.restype = C.c_int
    lib.ndp_status.argtypes = [vp, vp, vp]
    lib.ndp_stats.argtypes = [vp, vp, vp]
    lib.ndp_launch_count.argtypes = [vp]
    lib.ndp_launch_count.restype = i64
    lib.ndp_last_error.restype = C.c_char_p
    lib.ndp_rk4_sens.argtypes = [i32, i64, dbl, dbl, dbl, vp, vp, vp, vp, vp, vp]
    fp = C.POINTER(C.c_float)
    lib.ndp_mlp_create.argtypes = [fp] * 8 + [C.POINTER(vp)]
    lib.ndp_mlp_destroy.argtypes = [vp]
    lib.ndp_mlp_forward_pairs.argtypes = [vp, i32, i64, i32, vp, vp, vp, dbl, vp, i32, i32, vp]
    lib.ndp_mlp_forward_rows.argtypes = [vp, i64, vp, vp, i32, vp]
    lib.ndp_mlp_forward_swarm.argtypes = [vp, i32, i64, i64, i64, i32, vp, vp, dbl, vp, i32, vp]
    lib.ndp_mlp_launch_count.argtypes = [vp]
    lib.ndp_mlp_set_pair_budget.argtypes = [vp, i64]
    lib.ndp_mlp_set_pair_budget.restype = C.c_int
    lib.ndp_mlp_set_group.argtypes = [vp, i32]
    lib.ndp_mlp_set_group.restype = C.c_int
    lib.ndp_mlp_forward_swarm_parts.argtypes = [vp, i32, i32, C.POINTER(vp), i64, i64, i64, i64, i32, vp, dbl, vp, i32, vp]
    lib.ndp_mlp_forward_swarm_parts.restype = C.c_int
    lib.ndp_mlp_forward_pairs_ex.argtypes = [vp, i32, i64, i32, vp, vp, i32, vp, dbl, vp, i32, i32, vp]
    lib.ndp_pipeline_create.argtypes = [vp, vp, dbl, i32, C.POINTER(vp)]
    lib.ndp_pipeline_destroy.argtypes = [vp]
    lib.ndp_pipeline_buffers.argtypes = [vp, i32] + [C.POINTER(vp)] * 7
    lib.ndp_pipeline_submit.argtypes = [vp, i32]
    lib.ndp_pipeline_wait.argtypes = [vp, i32]
    lib.ndp_pipeline_bytes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.ndp_pipeline_stream.argtypes = [vp]
    lib.ndp_pipeline_stream.restype = vp
    lib.ndp_mlp_launch_count.restype = i64
    lib.ndp_plant_create.argtypes = [i64, dbl, dbl, i32, i32, i32, i64, C.POINTER(vp)]
    lib.ndp_plant_destroy.argtypes = [vp]
    lib.ndp_plant_reset.argtypes = [vp, vp]
    lib.ndp_plant_forward.argtypes = [vp, dbl, vp, vp, vp]
    lib.ndp_plant_autopilot.argtypes = [vp, vp, vp, dbl, vp]
    lib.ndp_plant_dynamics.argtypes = [vp, dbl, vp, vp]
    lib.ndp_plant_nmpc_x0.argtypes = [i64, vp, i32, vp, vp]
    lib.ndp_plant_cmd_from_u0.argtypes = [i64, i32, vp, dbl, dbl, vp, vp]
    lib.ndp_plant_launch_count.argtypes = [vp]
    lib.ndp_plant_launch_count.restype = i64
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.ndp_refgen_create.argtypes = [i32, ip, dp, dp, dp, dp, dp, dp, C.POINTER(vp)]
    lib.ndp_refgen_destroy.argtypes = [vp]
    lib.ndp_refgen_horizon.argtypes = [vp, i32, i64, vp, vp, i32, dbl, vp, vp, vp, vp]
    lib.ndp_refgen_launch_count.argtypes = [vp]
    lib.ndp_refgen_launch_count.restype = i64
    lib.ndp_predxu_len.argtypes = [i32]
    lib.ndp_predxu_len.restype = i64
    lib.ndp_predxu_pack.argtypes = [i32, i64, i32, vp, vp, vp, vp]
    lib.ndp_predxu_unpack.argtypes = [i32, i64, i32, vp, vp, vp, vp, vp]
    lib.ndp_hover_throttle_init.argtypes = [i64, vp, vp, vp]
    lib.ndp_hover_throttle_update.argtypes = [i64, dbl, vp, i64, vp, i64, vp, vp, vp]
    lib.ndp_plant_cmd_from_u0_dev.argtypes = [i64, i32, vp, dbl, vp, vp, vp]
    for name in ("ndp_predxu_pack", "ndp_predxu_unpack", "ndp_hover_throttle_init", "ndp_hover_throttle_update", "ndp_plant_cmd_from_u0_dev"):
        getattr(lib, name).restype = C.c_int
    lib.ndp_kernel_timing.argtypes = [vp, i32]
    lib.ndp_kernel_timing.restype = C.c_int
    lib.ndp_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.ndp_last_kernel_ms.restype = C.c_int
    lib.ndp_longlist_create.argtypes = [i32, i64, i32, i32, i32, i32, C.POINTER(vp)]
    lib.ndp_longlist_destroy.argtypes = [vp]
    lib.ndp_longlist_reset.argtypes = [vp, vp, vp, vp, vp]
    lib.ndp_longlist_push.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ndp_longlist_launch_count.argtypes = [vp]
    lib.ndp_longlist_launch_count.restype = i64
    lib.ndp_pipeline_create_ll.argtypes = [vp, vp, dbl, i32, vp, C.POINTER(vp)]
    for name in ("ndp_longlist_create", "ndp_longlist_destroy", "ndp_longlist_reset", "ndp_longlist_push", "ndp_pipeline_create_ll"):
        getattr(lib, name).restype = C.c_int
    for name in ("ndp_refgen_create", "ndp_refgen_destroy", "ndp_refgen_horizon"):
        getattr(lib, name).restype = C.c_int
    for name in ("ndp_plant_create", "ndp_plant_destroy", "ndp_plant_reset", "ndp_plant_forward", "ndp_plant_autopilot",
                 "ndp_plant_dynamics", "ndp_plant_nmpc_x0", "ndp_plant_cmd_from_u0"):
        getattr(lib, name).restype = C.c_int
    for name in ("ndp_create", "ndp_destroy", "ndp_set", "ndp_get", "ndp_reset", "ndp_set_reference", "ndp_solve", "ndp_update",
                 "ndp_status", "ndp_stats", "ndp_rk4_sens", "ndp_mlp_create", "ndp_mlp_destroy",
                 "ndp_mlp_forward_pairs", "ndp_mlp_forward_rows", "ndp_mlp_forward_swarm", "ndp_mlp_forward_pairs_ex",
                 "ndp_pipeline_create", "ndp_pipeline_destroy", "ndp_pipeline_buffers", "ndp_pipeline_submit", "ndp_pipeline_wait",
                 "ndp_pipeline_bytes"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ndp_last_error()
        raise NdpError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
