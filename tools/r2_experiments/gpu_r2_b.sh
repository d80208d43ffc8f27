# role timers of the persistent kernel (with and without the fused integrator) and producer-count variants
cd /root/repo
timeout 120 python -m pytest tests -m gpu -q -k "pair_list_molecular or state_invalidation" 2>&1 | tail -3
B="timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 1 --no-parity"
EMDEE_B200_LIB=/root/repo/build/libemdee_timing.so $B > gpurun_out/t_fused.json 2> gpurun_out/t_fused.err; grep "role timers" gpurun_out/t_fused.err | tail -2
EMDEE_FUSE_VV=0 EMDEE_B200_LIB=/root/repo/build/libemdee_timing.so $B > gpurun_out/t_unfused.json 2> gpurun_out/t_unfused.err; grep "role timers" gpurun_out/t_unfused.err | tail -2
bash tools/gpu_variants.sh np2 np3 2>&1 | tail -4
