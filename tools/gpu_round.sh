# One-GPU measurement pass of a round: tests, both bench arms, ncu launch list, ncu full capture of the top kernels.
cd /root/repo
R=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -2 gpurun_out/bench_$R.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; tail -2 gpurun_out/bench_ref_$R.err
timeout 600 python bench.py --workload c2 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_c2_$R.json 2> gpurun_out/bench_c2_$R.err
timeout 600 python bench.py --workload c4 --no-cpu-baseline --e2e-iters 1 > gpurun_out/bench_c4_$R.json 2> gpurun_out/bench_c4_$R.err
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-iters 1"
timeout 600 $B > gpurun_out/plain_$R.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$R.csv $B > gpurun_out/ncu_list_$R.log 2>&1
timeout 600 $B > gpurun_out/plain2_$R.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force_list_p|k_list_build' -s 6 -c 3 -o gpurun_out/prof_${R}_step $B > gpurun_out/ncu_full_$R.log 2>&1; tail -1 gpurun_out/ncu_full_$R.log
