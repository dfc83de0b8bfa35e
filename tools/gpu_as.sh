#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/swarm_diag.py --steps 12 > gpurun_out/swarm_diag.jsonl 2> gpurun_out/swarm_diag.err
timeout 400 python bench.py --no-cpu-baseline --no-latency > gpurun_out/bench.json 2> gpurun_out/bench.err
grep -a "as_first\|passed\|failed\|rc=" gpurun_out/pytest_gpu.log | tail -40; cat gpurun_out/swarm_diag.jsonl; tail -3 gpurun_out/swarm_diag.err; cut -c1-400 gpurun_out/bench.json
