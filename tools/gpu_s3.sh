mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/fin_smoke.log; tail -2 gpurun_out/fin_smoke.log
timeout 500 python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/fin_bench.json
