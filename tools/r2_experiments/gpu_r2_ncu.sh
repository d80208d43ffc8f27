# ncu: launch list and --set full capture of the stepping kernels (prune and replay launches) of the default bench command
cd /root/repo
R=${1:-r2b}
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-iters 1 --no-parity"
timeout 300 $B > gpurun_out/plain_$R.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv $B > gpurun_out/ncu_list_$R.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force_list_p' -s 12 -c 4 -o gpurun_out/prof_${R}_step $B > gpurun_out/ncu_full_$R.log 2>&1; tail -2 gpurun_out/ncu_full_$R.log
ls -la gpurun_out/prof_${R}_step.ncu-rep
