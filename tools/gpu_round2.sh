#!/bin/bash
# 2-GPU call: microbench, bench at N=1 and N=2 (both arms' launch path), coupled-swarm exchange check.
set -x
mkdir -p gpurun_out
./tools/microbench/ffma2 > gpurun_out/ffma2.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-latency > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/swarm_multi_gpu.py --quads 1024 --steps 50 --check --breakdown > gpurun_out/swarm_n2.json 2> gpurun_out/swarm_n2.err
timeout 300 python tools/swarm_multi_gpu.py --quads 1024 --steps 50 --check --breakdown > gpurun_out/swarm_n1.json 2> gpurun_out/swarm_n1.err
cat gpurun_out/ffma2.log; tail -c 600 gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err; cat gpurun_out/swarm_n2.json; tail -3 gpurun_out/swarm_n2.err; cat gpurun_out/swarm_n1.json
