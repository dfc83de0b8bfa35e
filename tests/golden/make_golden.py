"""Generate tests/golden/*.npz by running the REFERENCE's own code in this container.

Needs /root/reference (read-only mount); the GPU box does not have it, which is why the outputs are
committed.  What is generated, and from which reference code:
  mlp_golden.npz        nn_net.net (dnwash_nn_est/nn_net.py:7-18) + the shipped SN=4 checkpoint,
                        torch CPU fp32 and fp64, on fixed inputs (incl. SURVEY.md B.1 rows)
  downwash_golden.npz   DownwashNN.update semantics (downwash_nn.py:21-29) on two [21,10] horizons
  hv_throttle_golden.npz  HoverThrottleEstimator (hv_throttle_est/*.py) driven by a fixed sequence
  rti_golden.npz        NOT reference output (acados is absent): problems + solutions of the dense
                        numpy oracle, kept as a regression pin for the C oracle
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
REF = "/root/reference/ndp_nmpc/scripts"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


def mlp():
    from dnwash_nn_est.nn_net import net  # reference module

    net.load_state_dict(torch.load(os.path.join(REF, "dnwash_nn_est/nn_model/128-64-128_WBias_SN=4_epoch=20000_test_loss=1.0221.pkl"),
                                   map_location="cpu", weights_only=True))
    net.eval()
    rng = np.random.default_rng(0)
    x = np.concatenate([
        np.array([[0, 0, 0.5, 0, 0, 0], [0, 0, 1.0, 0, 0, 0], [0.2, -0.1, 0.7, 0.1, 0, -0.2], [0, 0, -0.5, 0, 0, 0],
                  [1, 1, 1.5, 0, 0, 0], [0, 0, 0, 0, 0, 0]], dtype=np.float64),
        rng.normal(size=(250, 6)) * np.array([1.0, 1.0, 1.5, 2.0, 2.0, 1.0]),
    ]).astype(np.float32)
    with torch.no_grad():
        y32 = net(torch.from_numpy(x)).numpy()
        net64 = net.double()
        y64 = net64(torch.from_numpy(x).double()).numpy()
        net.float()
    np.savez(os.path.join(OUT, "mlp_golden.npz"), x=x, y32=y32, y64=y64)
    print("mlp", y32[:2], np.abs(y32 - y64).max())
    # DownwashNN.update semantics without the hard-wired CUDA device: same arithmetic on CPU
    from ndp_nmpc_qd_b200 import workloads as wl

    ego, _ = wl.reference_horizon([3.0], name="eight_low")
    other = ego.copy()
    other[0, :, 0:3] += np.array([0.3, -0.2, 0.8])
    other[0, :, 3:6] += 0.05
    inp = (other[0] - ego[0])[:, 0:6]
    with torch.no_grad():
        f = net(torch.from_numpy(inp).to(torch.float32)).cpu().detach().numpy()
    np.savez(os.path.join(OUT, "downwash_golden.npz"), ego=ego[0], other=other[0], f=f)
    print("downwash", f[0])


def hv():
    from hv_throttle_est import HoverThrottleEstimator  # reference module

    est = HoverThrottleEstimator(0.02)
    rng = np.random.default_rng(1)
    n = 600
    thr = 0.27434003169930943 + 0.02 * np.sin(np.arange(n) * 0.05)
    thr[100:120] = 0.05  # below the 0.1 gate: update skipped
    vz = 0.3 * np.sin(np.arange(n) * 0.02) + 0.01 * rng.normal(size=n)
    k = np.zeros(n)
    xs = np.zeros((n, 2))
    for i in range(n):
        k[i], x, P = est.update(float(vz[i]), float(thr[i]))
        xs[i] = x[:, 0]
    np.savez(os.path.join(OUT, "hv_throttle_golden.npz"), vz=vz, thr=thr, k=k, x=xs, P_last=P)
    # SURVEY.md B.2 known answers
    est = HoverThrottleEstimator(0.02)
    kk = [est.update(0.0, 0.27434003169930943)[0] for _ in range(500)]
    print("hv first5", kk[:5], "after500", kk[-1])


def rti():
    from oracle import nmpc_numpy as on
    from ndp_nmpc_qd_b200 import workloads as wl

    p = on.OcpParams()
    recs = {k: [] for k in ("x0", "xr", "ur", "fd", "u0", "X", "U", "n_active", "scale")}
    for scale, seed in ((1.0, 11), (5.0, 12), (15.0, 13)):
        w = wl.independent_problems(4, seed=seed, scale=scale)
        fd = np.random.default_rng(seed).normal(size=(4, 21, 3)) * (scale > 1)
        for b in range(4):
            d = on.rti_step(w["x0"][b], w["xr"][b], w["ur"][b], fd[b], w["xr"][b].copy(), w["ur"][b].copy(), p, tol=1e-12)
            assert d["status"] == 0
            for k, v in (("x0", w["x0"][b]), ("xr", w["xr"][b]), ("ur", w["ur"][b]), ("fd", fd[b]), ("u0", d["u0"]), ("X", d["X"]),
                         ("U", d["U"]), ("n_active", d["n_active"]), ("scale", scale)):
                recs[k].append(v)
    np.savez(os.path.join(OUT, "rti_golden.npz"), **{k: np.array(v) for k, v in recs.items()})
    print("rti n_active", recs["n_active"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    mlp()
    hv()
    rti()
