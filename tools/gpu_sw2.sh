#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --check --breakdown --trace > gpurun_out/swarm_n2.json 2> gpurun_out/swarm_n2.err
timeout 300 python tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --check --breakdown > gpurun_out/swarm_n1.json 2> gpurun_out/swarm_n1.err
cat gpurun_out/swarm_n2.json; tail -3 gpurun_out/swarm_n2.err; cat gpurun_out/swarm_n1.json
