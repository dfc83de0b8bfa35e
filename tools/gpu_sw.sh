#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/swarm_diag.py --steps 12 > gpurun_out/swarm_diag.jsonl 2> gpurun_out/swarm_diag.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/swarm_launches.csv python tools/swarm_diag.py --steps 6 > gpurun_out/swarm_under_ncu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cut -c1-200 gpurun_out/swarm_diag.jsonl; tail -3 gpurun_out/swarm_diag.err
