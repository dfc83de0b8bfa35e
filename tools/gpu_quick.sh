#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.batch1_latency(torch.device('cuda', 0))))
PY
