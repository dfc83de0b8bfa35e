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


RTI_INST = os.path.join(HERE, "csrc", "rti_inst.cu")
# (tag, element type, horizon template argument (0 = run-time N), latency build)
RTI_INSTANCES = [("f32_20_0", "float", 20, "false"), ("f32_40_0", "float", 40, "false"), ("f32_80_0", "float", 80, "false"),
                 ("f32_0_0", "float", 0, "false"), ("f32_20_1", "float", 20, "true"),
                 ("f64_20_0", "double", 20, "false"), ("f64_40_0", "double", 40, "false"), ("f64_80_0", "double", 80, "false"),
                 ("f64_0_0", "double", 0, "false")]


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
    """Every translation unit is compiled with nvcc for sm_100a (the RTI kernel instantiations in parallel),
    then linked into one shared object.  defines / out: variant builds for A/B measurements (tools/)."""
    import concurrent.futures as cf

    if not force and out == OUT and not is_stale():
        return OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    obj_dir = os.path.join(OUT_DIR, "obj" if out == OUT else "obj_" + os.path.basename(out))
    os.makedirs(obj_dir, exist_ok=True)
    common = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v" if verbose else "-O3", "-ccbin", "/usr/bin/g++"] + [f"-D{d}" for d in defines]
    rdc = []
    jobs = [(os.path.join(obj_dir, "ndp_capi.o"), [SRC])]
    for tag, typ, n, lat in RTI_INSTANCES:
        jobs.append((os.path.join(obj_dir, f"rti_{tag}.o"),
                     [RTI_INST, f"-DNDP_INST_T={typ}", f"-DNDP_INST_N={n}", f"-DNDP_INST_LAT={lat}", f"-DNDP_INST_TAG={tag}",
                      f"-DNDP_INST_LAT_IS_TRUE={1 if lat == 'true' else 0}"] + rdc))

    def compile_one(job):
        obj, args = job
        return subprocess.run(common + ["-c", "-o", obj] + args, capture_output=True, text=True)

    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, jobs))
    for res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed building libndp_nmpc_b200.so")
        if verbose:
            sys.stderr.write(res.stderr)
    link = [_nvcc(), "-shared", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [j[0] for j in jobs] + ["-lcuda"]
    if rdc:
        link += ["-lcudadevrt"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libndp_nmpc_b200.so")
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
