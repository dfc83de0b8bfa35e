#!/bin/bash
# One gpurun call: GPU tests, both bench arms, ncu launch list, ncu --set full of the two hot kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rti_step_kernel -s 4 -c 2 -o gpurun_out/prof_rti -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/prof_rti.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 4 -c 2 -o gpurun_out/prof_mlp -f \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/prof_mlp.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json
timeout 900 python tools/closed_loop_sweep.py > gpurun_out/closed_loop_sweep.jsonl 2> gpurun_out/closed_loop_sweep.err
timeout 300 python tools/gpu_stress_sweep.py 0 6 10 16 > gpurun_out/stress_sweep.jsonl 2>&1
tail -2 gpurun_out/closed_loop_sweep.err; wc -l gpurun_out/closed_loop_sweep.jsonl
