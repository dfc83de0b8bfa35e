#!/bin/bash
# One gpurun call: GPU tests, smoke, the bench (own arm) — the "is the tree green" check.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json
