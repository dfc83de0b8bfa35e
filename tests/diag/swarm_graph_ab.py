import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
from ndp_nmpc_qd_b200.swarm import time_swarm
modes = ["p2p", "allgather"] if world > 1 else ["local"]
for g in (False, True):
    r = time_swarm(1024, modes, steps=100, warmup=60, use_graph=g)
    if not dist.is_initialized() or dist.get_rank() == 0:
        print("graph", g, json.dumps({k: (dict(ms=round(v["ms_per_step"] * 1e3, 1), mode=v["launch_mode"]) if isinstance(v, dict) else v) for k, v in r.items()}), flush=True)
if world > 1:
    dist.destroy_process_group()
