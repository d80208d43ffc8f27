# ncu --set full of one stepping launch of config 4 (compacted staging + split lists), after the same command exited 0 without ncu
cd /root/repo
B="python bench.py --workload c4 --steps 4 --warmup 2 --no-cpu-baseline --e2e-iters 1 --no-parity"
timeout 200 $B > gpurun_out/plain_c4ncu.log 2>&1 || { tail -3 gpurun_out/plain_c4ncu.log; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_force_list_p' -s 4 -c 1 -o gpurun_out/prof_r2_c4 $B > gpurun_out/ncu_c4.log 2>&1; tail -1 gpurun_out/ncu_c4.log | cut -c1-200
ls -la gpurun_out/prof_r2_c4.ncu-rep
