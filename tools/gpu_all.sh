#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/swarm_diag.py --steps 12 > gpurun_out/swarm_diag.jsonl 2> gpurun_out/swarm_diag.err
timeout 300 python tools/gpu_stress_sweep.py 0 6 10 > gpurun_out/stress_sweep.jsonl 2>&1
grep -a "warm active\|passed\|failed\|Error\|assert" gpurun_out/pytest_gpu.log | tail -12
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['roofline']['kernel_ms'], d['mlp']['kernel_ms'], d['e2e']['value'], json.dumps(d['stress']), json.dumps(d['latency_b1']))"
cut -c1-330 gpurun_out/swarm_diag.jsonl; tail -3 gpurun_out/swarm_diag.err; cat gpurun_out/stress_sweep.jsonl
