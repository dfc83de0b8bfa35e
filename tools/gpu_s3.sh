mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_pipeline.py tests/test_gpu_swarm_step.py -m gpu -x -q > gpurun_out/s3f_pytest.log 2>&1; tail -3 gpurun_out/s3f_pytest.log
for rep in 1 2; do
NDP_NMPC_LIB=ndp_nmpc_qd_b200/_C/variants/lib_l4smem.so timeout 200 python bench.py --kernels-only --steps 40 --warmup 5 2>/dev/null | tail -1 | cut -c1-400
timeout 200 python bench.py --kernels-only --steps 40 --warmup 5 2>/dev/null | tail -1 | cut -c1-400
done
