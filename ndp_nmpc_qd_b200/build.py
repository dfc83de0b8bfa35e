"""Builds libndp_nmpc_b200.so in-tree with nvcc for sm_100a (no torch extension machinery:
the library is a plain C-ABI shared object, see include/ndp_nmpc.h)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "ndp_capi.cu")
OUT_DIR = os.path.join(HERE, "_C")
OUT = os.path.join(OUT_DIR, "libndp_nmpc_b200.so")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or install the CUDA toolkit)")


def sources():
    d = os.path.join(HERE, "csrc")
    inc = os.path.join(os.path.dirname(HERE), "include", "ndp_nmpc.h")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [inc]


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False, defines=(), out: str = OUT) -> str:
    """defines / out: variant builds for A/B measurements (tools/), the product is the default."""
    if not force and out == OUT and not is_stale():
        return OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
        "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v" if verbose else "-O3",
        "-ccbin", "/usr/bin/g++", "-o", out, SRC, "-lcuda",
    ] + [f"-D{d}" for d in defines]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libndp_nmpc_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
