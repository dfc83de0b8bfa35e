#!/bin/bash
# 8-GPU call on the final tree: bench at N=8 (its line carries the swarm and closed-loop blocks at 8 ranks).
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/fin_bench_n8.json 2> gpurun_out/fin_bench_n8.err
cut -c1-260 gpurun_out/fin_bench_n8.json; tail -2 gpurun_out/fin_bench_n8.err
