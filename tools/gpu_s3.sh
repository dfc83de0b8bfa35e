mkdir -p gpurun_out
for v in "" ndp_nmpc_qd_b200/_C/variants/lib_cta256.so ndp_nmpc_qd_b200/_C/variants/lib_cta128.so; do
  echo "== lib ${v:-default (1024)}"
  NDP_NMPC_LIB=$v timeout 200 python tests/diag/swarm_multi_gpu.py --quads 1024 --steps 200 --check --breakdown 2>/dev/null | cut -c1-700
done
NDP_NMPC_LIB=ndp_nmpc_qd_b200/_C/variants/lib_cta256.so timeout 200 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_swarm_step.py -m gpu -x -q 2>&1 | tail -2
