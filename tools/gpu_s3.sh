mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/s3e_pytest.log 2>&1; tail -3 gpurun_out/s3e_pytest.log
NDP_NMPC_LIB=ndp_nmpc_qd_b200/_C/variants/lib_base.so timeout 120 python tests/diag/gpu_diag_ab_identical.py dump /tmp/a.npz 2>&1 | tail -3
timeout 120 python tests/diag/gpu_diag_ab_identical.py dump /tmp/b.npz 2>&1 | tail -3
python tests/diag/gpu_diag_ab_identical.py cmp /tmp/a.npz /tmp/b.npz > gpurun_out/s3e_identical.log 2>&1; cat gpurun_out/s3e_identical.log
timeout 150 python tools/gpu_stress_sweep.py 6 7 8 10 > gpurun_out/s3e_stress.jsonl 2>&1; cut -c1-200 gpurun_out/s3e_stress.jsonl
for v in base prof; do :; done
NDP_NMPC_LIB=ndp_nmpc_qd_b200/_C/variants/lib_base.so timeout 200 python bench.py --kernels-only --steps 40 --warmup 5 2>/dev/null | tail -1 | cut -c1-400
timeout 200 python bench.py --kernels-only --steps 40 --warmup 5 2>/dev/null | tail -1 | cut -c1-400
