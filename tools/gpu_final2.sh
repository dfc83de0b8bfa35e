#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/closed_loop_sweep.py > gpurun_out/closed_loop_sweep.jsonl 2> gpurun_out/closed_loop_sweep.err
bash tools/gpu_round2.sh
head -2 gpurun_out/closed_loop_sweep.jsonl | cut -c1-200
