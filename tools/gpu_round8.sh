#!/bin/bash
# 8-GPU call: bench (weak scaling, no collective) and the coupled swarm (config 4) at 8 and 4 GPUs.
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --check --breakdown --trace > gpurun_out/swarm_n8.json 2> gpurun_out/swarm_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --check --breakdown --trace > gpurun_out/swarm_n4.json 2> gpurun_out/swarm_n4.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 tests/diag/swarm_multi_gpu.py --quads 8192 --steps 100 --breakdown --trace > gpurun_out/swarm_n8_8192.json 2> gpurun_out/swarm_n8_8192.err
tail -c 700 gpurun_out/bench_n8.json; tail -2 gpurun_out/bench_n8.err; cat gpurun_out/swarm_n8.json; tail -3 gpurun_out/swarm_n8.err; cat gpurun_out/swarm_n4.json; cat gpurun_out/swarm_n8_8192.json; tail -3 gpurun_out/swarm_n8_8192.err
