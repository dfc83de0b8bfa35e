"""GPU parity tests of the batched dop_sim plant: CUDA kernels through the C ABI vs the
reference's own TorchScript module (golden fixture) and vs the numpy oracle on larger random swarms."""
import numpy as np
import pytest

from conftest import golden
from oracle.plant_numpy import PlantOracle

pytestmark = pytest.mark.gpu

CASES = {"b4": (0.01, 0.01, (True, True, True)), "rnd": (0.01, 0.02, (True, True, True)), "ind": (0.01, 0.02, (False, True, False))}
TOL = 1e-9  # float64 kernels; FMA contraction / libm differences only


def _err(got, ref):
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    g, r = np.nan_to_num(got), np.nan_to_num(ref)
    return float((np.abs(g - r) / np.maximum(np.abs(r), 1.0)).max())


def _gpu_rollout(n, ts_sim, ts_ctl, flags, s0, cmds, idx, group=0):
    import torch
    from ndp_nmpc_qd_b200.dop_sim import MulQuadrotors

    m = MulQuadrotors(n, ts_sim, ts_ctl, torch.float64, *flags, group=group)
    s = torch.as_tensor(s0.copy(), device="cuda").reshape(n, 35, 1).contiguous()
    out = []
    for k, c in enumerate(cmds):
        r = m(ts_sim, s, torch.as_tensor(c.copy(), device="cuda").reshape(n, 4, 1).contiguous())
        assert r.data_ptr() == s.data_ptr()  # in place, like the reference
        if k in idx:
            out.append(s.cpu().numpy()[:, :, 0].copy())
    return np.stack(out)


@pytest.mark.parametrize("case", list(CASES))
def test_plant_matches_reference_module(built_lib, case):
    g = golden("plant_golden.npz")
    ts_sim, ts_ctl, flags = CASES[case]
    s0, cmds, idx, ref = g[case + "_s0"], g[case + "_cmd"], set(g[case + "_idx"].tolist()), g[case + "_states"]
    got = _gpu_rollout(s0.shape[0], ts_sim, ts_ctl, flags, s0, cmds, idx)
    assert _err(got, ref) < TOL


def _random_swarm(n, steps, seed):
    rng = np.random.default_rng(seed)
    s0 = np.zeros((n, 35))
    side = max(2.0, (n / 8.0) ** (1 / 3))  # ~8 quads per cubic metre column: plenty of downwash pairs
    s0[:, 3:6] = rng.uniform(0, side, size=(n, 3))
    q = rng.normal(size=(n, 4)) * np.array([1, 0.2, 0.2, 0.2]); q[:, 0] = np.abs(q[:, 0]) + 1
    s0[:, 9:13] = q / np.linalg.norm(q, axis=1, keepdims=True)
    s0[:, 13:16] = rng.normal(size=(n, 3)); s0[:, 16:19] = rng.normal(size=(n, 3)) * 0.5
    s0[:, 19:22] = rng.normal(size=(n, 3)) * 0.3; s0[:, 28:31] = rng.normal(size=(n, 3)) * 0.2
    s0[:, 31:35] = 8 + rng.normal(size=(n, 4))
    cmds = np.zeros((steps, n, 4))
    cmds[:, :, 0:3] = rng.normal(size=(steps, n, 3)) * 0.5
    cmds[:, :, 3] = np.clip(0.283 + 0.1 * rng.normal(size=(steps, n)), 0.0, 1.0)
    return s0, cmds


@pytest.mark.parametrize("n,group", [(1000, 0), (777, 0), (4096, 16)])
def test_plant_vs_oracle_large(built_lib, n, group):
    """ragged sizes (not a multiple of the CTA), all-pairs downwash and block-diagonal scenario groups."""
    steps = 12
    s0, cmds = _random_swarm(n, steps, seed=n)
    idx = {0, 5, steps - 1}
    got = _gpu_rollout(n, 0.01, 0.02, (True, True, True), s0, cmds, idx, group=group)
    if group:  # oracle: independent groups
        ref = np.zeros_like(got)
        for g0 in range(0, n, group):
            sl = slice(g0, g0 + group)
            o = PlantOracle(group, 0.01, 0.02, True, True, True)
            s, out = s0[sl].copy(), []
            for k, c in enumerate(cmds):
                s = o.forward(0.01, s, c[sl])
                if k in idx:
                    out.append(s.copy())
            ref[:, sl] = np.stack(out)
    else:
        o = PlantOracle(n, 0.01, 0.02, True, True, True)
        s, out = s0.copy(), []
        for k, c in enumerate(cmds):
            s = o.forward(0.01, s, c)
            if k in idx:
                out.append(s.copy())
        ref = np.stack(out)
    assert _err(got, ref) < TOL


def test_nmpc_glue_kernels(built_lib):
    import torch
    from ndp_nmpc_qd_b200.dop_sim import MulQuadrotors

    n = 300
    s0, _ = _random_swarm(n, 1, seed=2)
    m = MulQuadrotors(n, 0.01, 0.02, has_downwash=False)
    s = torch.as_tensor(s0, device="cuda").reshape(n, 35, 1).contiguous()
    for dt in (torch.float32, torch.float64):
        x0 = m.nmpc_x0(s, torch.empty((n, 10), dtype=dt, device="cuda")).cpu().numpy()
        ref = np.concatenate([s0[:, 3:6], s0[:, 13:16], s0[:, 9:13]], 1)  # pt_publisher.py:106-122
        assert np.array_equal(x0, ref.astype(x0.dtype))
        u0 = torch.as_tensor(np.random.default_rng(0).normal(size=(n, 4)), dtype=dt, device="cuda")
        cmd = m.cmd_from_u0(u0, torch.empty((n, 4, 1), dtype=torch.float64, device="cuda"), 1.4844, 50.0).cpu().numpy()[:, :, 0]
        u = u0.cpu().numpy().astype(np.float64)
        assert np.array_equal(cmd[:, 0:3], u[:, 0:3]) and np.allclose(cmd[:, 3], u[:, 3] * 1.4844 / 50.0, rtol=1e-15)
