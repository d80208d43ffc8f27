# Closing pass of a session (one GPU): every GPU test, the default bench, config 4 at both cell geometries, and an
# ncu --set full capture of config 4's stepping kernel (after the same command ran clean without ncu).
cd /root/repo
R=${1:-r1i}
timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/tests_$R.log; tail -4 gpurun_out/tests_$R.log
timeout 120 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -2 gpurun_out/bench_$R.err; cut -c1-330 gpurun_out/bench_$R.json
C4="python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --e2e-iters 1"
EMDEE_DEBUG=1 timeout 100 $C4 --ndiv 2 > gpurun_out/bench_c4n2_$R.json 2> gpurun_out/bench_c4n2_$R.err; grep "bricks\|pair list" gpurun_out/bench_c4n2_$R.err | tail -3; cut -c1-330 gpurun_out/bench_c4n2_$R.json
timeout 100 $C4 > gpurun_out/plain_c4_$R.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_force_list|k_list_build' -s 4 -c 2 -o gpurun_out/prof_${R}_c4 $C4 > gpurun_out/ncu_c4_$R.log 2>&1; tail -1 gpurun_out/ncu_c4_$R.log
