cd /root/repo
mkdir -p gpurun_out
bash tools/gpu_multi.sh 2
timeout 200 python -m pytest tests -m gpu -q -k "state_invalidation or cells_bit_exact" 2>&1 | tail -3
bash tools/gpu_variants.sh c16i2 c16i4 2>&1 | tail -3
