#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants2.txt
for f in ndp_nmpc_qd_b200/_C/variants/lib_*.so; do
  cp $f ndp_nmpc_qd_b200/_C/libndp_nmpc_b200.so
  echo $f >> gpurun_out/variants2.txt
  timeout 200 python tools/swarm_diag.py --steps 8 2>&1 | cut -c1-110 | tail -5 >> gpurun_out/variants2.txt
  timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_v.json')); print('rti_ms', round(d['roofline']['kernel_ms'],5), json.dumps(d['latency_b1'])[:120])" >> gpurun_out/variants2.txt
done
cat gpurun_out/variants2.txt
