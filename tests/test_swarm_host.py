"""CPU tests of the coupled-swarm host logic: shard bounds, and a world-size-2 gloo run of the exchange
(all_gather of the packed horizons) whose gathered tensor reproduces the single-process ordering."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from ndp_nmpc_qd_b200.swarm import shard_bounds


def test_shard_bounds_cover_and_pad():
    for n_all, world in [(1024, 8), (1000, 8), (7, 4), (3, 8), (1, 1)]:
        parts = [shard_bounds(n_all, world, r) for r in range(world)]
        part = parts[0][0]
        assert all(p[0] == part for p in parts) and part * world >= n_all
        covered = []
        for _, b, e in parts:
            assert 0 <= b <= e <= n_all and e - b <= part
            covered += list(range(b, e))
        assert covered == list(range(n_all))


WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from ndp_nmpc_qd_b200.swarm import shard_bounds, pack_horizons
from oracle import mlp_numpy
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_all, N = 37, 20
rng = np.random.default_rng(0)
xr_all = rng.normal(size=(n_all, N + 1, 10))
part, b, e = shard_bounds(n_all, world, rank)
mine = torch.zeros((part, N + 1, 6), dtype=torch.float32)
pack_horizons(torch.as_tensor(xr_all[b:e]), mine[: e - b])
gathered = torch.zeros((world * part, N + 1, 6), dtype=torch.float32)
dist.all_gather_into_tensor(gathered, mine)
ref = xr_all[:, :, 0:6].astype(np.float32)
assert np.array_equal(gathered[:n_all].numpy(), ref), "gathered ordering"
# the local egos' forces from the gathered tensor == the single-process oracle on the full swarm
w = mlp_numpy.load_npz(os.path.join(sys.argv[1], "ndp_nmpc_qd_b200/dnwash_nn_est/nn_model/128-64-128_WBias_SN=4_epoch=20000_test_loss=1.0221.npz"))
full = mlp_numpy.swarm_forces(w, ref, 0, n_all)
loc = mlp_numpy.swarm_forces(w, gathered[:n_all].numpy(), b, e - b)
assert np.array_equal(loc, full[b:e])
dist.barrier()
print("rank", rank, "ok")
"""


def test_gloo_world2_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(script), ROOT], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
