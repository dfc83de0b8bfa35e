"""Generate tests/golden/plant_golden.npz by running the REFERENCE's own TorchScript plant
(/root/reference/dop_sim/scripts/quadrotor, MulQuadrotors, float64, CPU) in this container.
The GPU box has no /root/reference, which is why the outputs are committed.

cases:
  b4_*    3 quads, ts_sim = ts_ctl = 0.01, downwash + motor + battery, 100 steps
  rnd_*   24 quads in a 3 m cube (many downwash pairs), random attitudes / rates / wind / body velocities,
          time-varying body-rate commands, ts_sim 0.01, ts_ctl 0.02, 60 steps (states after every step)
  ind_*   same without downwash (config-5 style independent scenarios), no battery
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, "/root/reference/dop_sim/scripts")
from quadrotor.mul_quadrotors import MulQuadrotors  # noqa: E402  (reference module)


def run(n, ts_sim, ts_ctl, flags, s0, cmds):
    m = torch.jit.script(MulQuadrotors(n, ts_sim, ts_ctl, torch.float64, *flags).requires_grad_(False))
    s = torch.from_numpy(s0.copy()).unsqueeze(2)
    out = []
    for c in cmds:
        s = m(ts_sim, s, torch.from_numpy(c.copy()).unsqueeze(2))
        out.append(s[:, :, 0].numpy().copy())
    return np.stack(out)


def random_case(seed, n, steps):
    rng = np.random.default_rng(seed)
    s0 = np.zeros((n, 35))
    s0[:, 3:6] = rng.uniform(0, 3, size=(n, 3)) + np.array([0, 0, 1.0])
    q = rng.normal(size=(n, 4)) * np.array([1, 0.2, 0.2, 0.2]); q[:, 0] = np.abs(q[:, 0]) + 1
    s0[:, 9:13] = q / np.linalg.norm(q, axis=1, keepdims=True)
    s0[: n // 4, 9:13] *= -1  # PX4-style negative scalar part
    s0[:, 13:16] = rng.normal(size=(n, 3))
    s0[:, 16:19] = rng.normal(size=(n, 3)) * 0.5  # body velocities (never integrated by the reference, feed the drag)
    s0[: n // 3, 16] = 0.0                        # u_r == 0 branch of alpha
    s0[:, 19:22] = rng.normal(size=(n, 3)) * 0.3
    s0[:, 28:31] = rng.normal(size=(n, 3)) * 0.2  # wind
    s0[: n // 3, 28:31] = 0.0
    s0[:, 31:35] = 8 + rng.normal(size=(n, 4))
    cmds = np.zeros((steps, n, 4))
    cmds[:, :, 0:3] = rng.normal(size=(steps, n, 3)) * 0.5
    cmds[:, :, 3] = np.clip(0.283 + 0.1 * rng.normal(size=(steps, n)), 0.0, 1.0)
    cmds[steps // 2:, : n // 6, 3] = 0.02  # thrust command clamps to zero
    return s0, cmds


def main():
    out = {}
    s0 = np.zeros((3, 35))
    s0[:, 3:6] = np.array([[1, 1, .5], [1, 2, .5], [1, 1.2, 1.5]], dtype=np.float32)
    s0[:, 9] = -1.0; s0[:, 31:35] = 8
    cmd = np.zeros((100, 3, 4)); cmd[:, :, 3] = 0.283
    out["b4_s0"], out["b4_cmd"] = s0, cmd
    out["b4_states"] = run(3, 0.01, 0.01, (True, True, True), s0, cmd)
    print("B.4 z", out["b4_states"][-1][:, 5], "vz", out["b4_states"][-1][:, 15], "o1", out["b4_states"][-1][:, 31])
    s0, cmd = random_case(0, 24, 60)
    out["rnd_s0"], out["rnd_cmd"] = s0, cmd
    out["rnd_states"] = run(24, 0.01, 0.02, (True, True, True), s0, cmd)
    s0, cmd = random_case(1, 16, 40)
    out["ind_s0"], out["ind_cmd"] = s0, cmd
    out["ind_states"] = run(16, 0.01, 0.02, (False, True, False), s0, cmd)
    # keep the fixture small: states after selected steps only (idx = step indices, 0-based)
    for name, idx in (("b4", [0, 1, 2, 49, 99]), ("rnd", list(range(0, 60, 6)) + [59]), ("ind", list(range(0, 40, 8)) + [39])):
        out[name + "_idx"] = np.array(idx)
        out[name + "_states"] = out[name + "_states"][idx]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "plant_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
