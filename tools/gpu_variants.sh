#!/bin/bash
# A/B of build variants (ndp_nmpc_qd_b200/_C/variants/lib_*.so): nominal bench kernel time, swarm tail, stress
mkdir -p gpurun_out
: > gpurun_out/variants.txt
for f in ndp_nmpc_qd_b200/_C/variants/lib_*.so; do
  cp $f ndp_nmpc_qd_b200/_C/libndp_nmpc_b200.so
  for rep in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline --no-latency > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_v.json')); print('$f', 'value', round(d['value']/1e6,3), 'rti_ms', round(d['roofline']['kernel_ms'],5), 'mlp_ms', round(d['mlp']['kernel_ms'],5))" >> gpurun_out/variants.txt
  done
  timeout 200 python tools/gpu_stress_sweep.py 6 2>&1 | cut -c1-120 >> gpurun_out/variants.txt
done
cat gpurun_out/variants.txt
