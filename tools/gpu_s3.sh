mkdir -p gpurun_out
for v in ndp_nmpc_qd_b200/_C/variants/lib_base.so ""; do
  echo "== lib ${v:-new}"
  NDP_NMPC_LIB=$v timeout 200 python bench.py --kernels-only --steps 40 --warmup 5 2>/dev/null | tail -1 | cut -c1-300
  NDP_NMPC_LIB=$v timeout 200 python tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --breakdown 2>/dev/null | cut -c60-420
  NDP_NMPC_LIB=$v timeout 100 python tests/diag/gpu_diag_lat.py 2>/dev/null | tail -3
done
NDP_NMPC_LIB=ndp_nmpc_qd_b200/_C/variants/lib_base.so timeout 120 python tests/diag/gpu_diag_ab_identical.py dump /tmp/a.npz 2>&1 | tail -3
timeout 120 python tests/diag/gpu_diag_ab_identical.py dump /tmp/b.npz 2>&1 | tail -3
python tests/diag/gpu_diag_ab_identical.py cmp /tmp/a.npz /tmp/b.npz 2>&1 | grep -c "identical True"
