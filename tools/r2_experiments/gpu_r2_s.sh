cd /root/repo
B="timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 1 --no-parity"
for tma in 0 1; do
EMDEE_TMA=$tma EMDEE_B200_LIB=/root/repo/build/libemdee_timing.so $B > gpurun_out/t_tma$tma.json 2> gpurun_out/t_tma$tma.err; echo "TMA=$tma fused: $(grep 'role timers' gpurun_out/t_tma$tma.err | tail -1)"
EMDEE_FUSE_VV=0 EMDEE_TMA=$tma EMDEE_B200_LIB=/root/repo/build/libemdee_timing.so $B > gpurun_out/t_tma${tma}u.json 2> gpurun_out/t_tma${tma}u.err; echo "TMA=$tma unfused: $(grep 'role timers' gpurun_out/t_tma${tma}u.err | tail -1)"
done
