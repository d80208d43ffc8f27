# Round 2 closing pass on one GPU (second half of the round: mbarrier hand-over, compute_nonbonded_into, two-level list):
# checked build over the GPU tests, the GPU tests, both bench arms, configs 2 / 4 / 5 on one GPU, a 1000-step run, the ncu launch
# list of the default bench command and a --set full capture of the step kernels (each ncu run after the same command exited 0).
cd /root/repo
R=${1:-r2c}
mkdir -p gpurun_out
bash tools/gpu_checked.sh
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/tests_$R.log; tail -3 gpurun_out/tests_$R.log
timeout 300 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -2 gpurun_out/bench_$R.err; cut -c1-300 gpurun_out/bench_$R.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; cut -c1-200 gpurun_out/bench_ref_$R.json
timeout 200 python bench.py --workload c2 --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_c2_$R.json 2> gpurun_out/bench_c2_$R.err; cut -c1-260 gpurun_out/bench_c2_$R.json
timeout 200 python bench.py --workload c4 --no-cpu-baseline --e2e-iters 1 > gpurun_out/bench_c4_$R.json 2> gpurun_out/bench_c4_$R.err; cut -c1-260 gpurun_out/bench_c4_$R.json
timeout 200 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --e2e-iters 1 > gpurun_out/bench_c3_long_$R.json 2> gpurun_out/bench_c3_long_$R.err; cut -c1-260 gpurun_out/bench_c3_long_$R.json
timeout 280 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --e2e-iters 1 > gpurun_out/bench_c5_$R.json 2> gpurun_out/bench_c5_$R.err; cut -c1-260 gpurun_out/bench_c5_$R.json
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-iters 1 --no-parity"
timeout 200 $B > gpurun_out/plain_$R.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$R.csv $B > gpurun_out/ncu_list_$R.log 2>&1; tail -1 gpurun_out/ncu_list_$R.log | cut -c1-200
timeout 200 $B > gpurun_out/plain2_$R.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_force_list_p|k_list_build' -s 6 -c 3 -o gpurun_out/prof_${R}_step $B > gpurun_out/ncu_full_$R.log 2>&1; tail -1 gpurun_out/ncu_full_$R.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
