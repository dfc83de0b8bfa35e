#!/bin/bash
# 2-GPU call on the final tree: the multi-GPU test, bench at N=2, coupled swarm in both exchange modes.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_swarm_step.py -m gpu -x -q > gpurun_out/fin2_pytest.log 2>&1; tail -2 gpurun_out/fin2_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/fin_bench_n2.json 2> gpurun_out/fin_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --check --breakdown > gpurun_out/fin_swarm_n2.json 2> gpurun_out/fin_swarm_n2.err
cut -c1-300 gpurun_out/fin_bench_n2.json; tail -2 gpurun_out/fin_bench_n2.err; cat gpurun_out/fin_swarm_n2.json | cut -c1-900
