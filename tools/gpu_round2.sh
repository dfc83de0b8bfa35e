#!/bin/bash
# 2-GPU call: bench at N=2 (weak scaling, no collective), coupled swarm (config 4) at 2 and 1 GPUs with the
# per-phase device times, the dense un-gated swarm case.
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --check --breakdown --trace > gpurun_out/swarm_n2.json 2> gpurun_out/swarm_n2.err
timeout 300 python tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --check --breakdown --trace > gpurun_out/swarm_n1.json 2> gpurun_out/swarm_n1.err
timeout 300 python tools/swarm_diag.py --steps 4 --dense > gpurun_out/swarm_dense.jsonl 2> gpurun_out/swarm_dense.err
tail -c 600 gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err; cat gpurun_out/swarm_n2.json; tail -3 gpurun_out/swarm_n2.err; cat gpurun_out/swarm_n1.json; tail -1 gpurun_out/swarm_dense.jsonl
