mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/s3c_pytest.log 2>&1; tail -3 gpurun_out/s3c_pytest.log
NDP_NMPC_LIB=ndp_nmpc_qd_b200/_C/variants/lib_base.so timeout 120 python tests/diag/gpu_diag_ab_identical.py dump /tmp/a.npz 2>&1 | tail -3
timeout 120 python tests/diag/gpu_diag_ab_identical.py dump /tmp/b.npz 2>&1 | tail -3
python tests/diag/gpu_diag_ab_identical.py cmp /tmp/a.npz /tmp/b.npz > gpurun_out/s3c_identical.log 2>&1; cat gpurun_out/s3c_identical.log
NDP_NMPC_LIB=ndp_nmpc_qd_b200/_C/variants/lib_prof.so timeout 120 python tests/diag/gpu_diag_lone_ipm.py > gpurun_out/s3c_lone.log 2>&1; tail -16 gpurun_out/s3c_lone.log
timeout 150 python tools/gpu_stress_sweep.py 6 7 8 10 > gpurun_out/s3c_stress.jsonl 2>&1; cut -c1-200 gpurun_out/s3c_stress.jsonl
