"""A/B measurement of kernel variants: builds libndp_nmpc_b200.so with extra -D flags into ndp_nmpc_qd_b200/_C/variants/
(here, on the CPU box) or runs `bench.py --kernels-only` against each of them (on the GPU box).
  python tools/ab_variants.py build NAME=FLAG1,FLAG2 ...      # NAME= (no flags) is the baseline
  python tools/ab_variants.py run [bench args]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "ndp_nmpc_qd_b200", "_C", "variants")

if sys.argv[1] == "build":
    from ndp_nmpc_qd_b200 import build

    os.makedirs(VDIR, exist_ok=True)
    for spec in sys.argv[2:]:
        name, _, flags = spec.partition("=")
        out = build.build(force=True, defines=[f for f in flags.split(",") if f], out=os.path.join(VDIR, f"lib_{name}.so"))
        print(name, out)
else:
    for f in sorted(os.listdir(VDIR)):
        if not f.endswith(".so"):
            continue
        env = dict(os.environ, NDP_NMPC_LIB=os.path.join(VDIR, f))
        res = []
        for rep in range(2):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--kernels-only", "--steps", "40", "--warmup", "5"] + sys.argv[2:],
                               env=env, capture_output=True, text=True)
            try:
                res.append(json.loads(r.stdout.strip().splitlines()[-1]))
            except Exception:
                res.append(dict(error=r.stderr[-300:]))
        print(f, json.dumps(res), flush=True)
