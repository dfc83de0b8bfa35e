#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/swarm_diag.py --steps 12 > gpurun_out/swarm_diag.jsonl 2> gpurun_out/swarm_diag.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:rti_step_kernel -s 8 -c 1 -o gpurun_out/prof_rti_swarm -f \
  python tools/swarm_diag.py --steps 10 > gpurun_out/prof_rti_swarm.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rti_step_kernel -s 4 -c 1 -o gpurun_out/prof_rti -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/prof_rti.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 4 -c 1 -o gpurun_out/prof_mlp -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/prof_mlp.log 2>&1
cat gpurun_out/swarm_diag.jsonl; tail -3 gpurun_out/swarm_diag.err
