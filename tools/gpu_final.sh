#!/bin/bash
# One gpurun call on the final tree: GPU tests, smoke, both bench arms (timed), ncu launch list, bench --profile.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/fin_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - t0 )) s" >> gpurun_out/fin_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/fin_smoke.log
t1=$(date +%s)
timeout 500 python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$? $(( $(date +%s) - t1 )) s" >> gpurun_out/fin_bench.err
t1=$(date +%s)
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_bench_ref.json 2> gpurun_out/fin_bench_ref.err; echo "ref rc=$? $(( $(date +%s) - t1 )) s" >> gpurun_out/fin_bench_ref.err
timeout 300 python bench.py --dtype f64 --no-cpu-baseline > gpurun_out/fin_bench_f64.json 2> gpurun_out/fin_bench_f64.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/fin_bench_under_ncu.log 2>&1
timeout 400 python bench.py --profile > gpurun_out/fin_profile.json 2> gpurun_out/fin_profile.err; cp profiles/ncu_rti_summary.json gpurun_out/fin_ncu_rti_summary.json
tail -2 gpurun_out/fin_pytest.log; tail -2 gpurun_out/fin_smoke.log; cut -c1-700 gpurun_out/fin_bench.json; tail -1 gpurun_out/fin_bench.err; tail -1 gpurun_out/fin_bench_ref.err; cut -c1-300 gpurun_out/fin_bench_ref.json
