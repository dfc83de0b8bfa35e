mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/s3g_pytest.log 2>&1; tail -2 gpurun_out/s3g_pytest.log
NDP_NMPC_LIB=ndp_nmpc_qd_b200/_C/variants/lib_base.so timeout 120 python tests/diag/gpu_diag_ab_identical.py dump /tmp/a.npz 2>&1 | tail -3
timeout 120 python tests/diag/gpu_diag_ab_identical.py dump /tmp/b.npz 2>&1 | tail -3
python tests/diag/gpu_diag_ab_identical.py cmp /tmp/a.npz /tmp/b.npz > gpurun_out/s3g_identical.log 2>&1; grep -c "identical True" gpurun_out/s3g_identical.log; grep "identical False" gpurun_out/s3g_identical.log
NDP_NMPC_LIB=ndp_nmpc_qd_b200/_C/variants/lib_prof.so timeout 120 python tests/diag/gpu_diag_lone_ipm.py > gpurun_out/s3g_lone.log 2>&1; tail -16 gpurun_out/s3g_lone.log
for v in ndp_nmpc_qd_b200/_C/variants/lib_base.so ""; do
  echo "== lib ${v:-new}"
  NDP_NMPC_LIB=$v timeout 150 python tools/gpu_stress_sweep.py 8 2>&1 | cut -c1-120
  NDP_NMPC_LIB=$v timeout 200 python tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --breakdown 2>/dev/null | cut -c60-420
done
