"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/ndp_nmpc.h declares (no compute calls without a GPU), and the product path refuses to run
without CUDA instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from ndp_nmpc_qd_b200 import _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "ndp_nmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ndp_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    for name in header_functions():
        assert hasattr(lib, name), name


def test_default_config_matches_reference_constants(built_lib):
    lib = _lib.load()
    cfg = _lib.NdpConfig()
    lib.ndp_default_config(C.byref(cfg))
    # params/nmpc_params.py:9-35, params/fhnp_params.py:9-19
    assert cfg.N == 20 and cfg.T == 2.0 and cfg.np == 4
    assert list(cfg.Q) == [300, 300, 400, 10, 10, 10, 0, 10, 10, 100]
    assert list(cfg.R) == [10, 10, 10, 5]
    assert list(cfg.u_min) == [-6, -6, -6, 0]
    assert abs(cfg.u_max[3] - 9.81 / 0.36) < 1e-12 and list(cfg.u_max)[:3] == [6, 6, 6]
    assert list(cfg.v_min) == [-20] * 3 and list(cfg.v_max) == [20] * 3
    assert cfg.mass == 1.4844 and cfg.gravity == 9.81


def test_api_misuse_is_reported(built_lib):
    lib = _lib.load()
    assert lib.ndp_create(None, None) == -1
    assert b"null" in lib.ndp_last_error()
    assert lib.ndp_solve(None, None, None, None) == -1
    assert lib.ndp_destroy(None) == 0


def test_sass_is_sm100a(built_lib):
    import subprocess

    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(built_lib):
    from ndp_nmpc_qd_b200.dnwash_nn_est import DownwashNN
    from ndp_nmpc_qd_b200.nmpc_ctl import NMPCBodyRateController

    with pytest.raises(_lib.NdpError):
        NMPCBodyRateController()
    with pytest.raises(_lib.NdpError):
        DownwashNN()
