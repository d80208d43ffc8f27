# Short validation pass (one GPU): all GPU tests (no -x: every failure is listed), the default bench, config 4, a long run.
cd /root/repo
R=${1:-r1g}
timeout 280 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/tests_$R.log; tail -15 gpurun_out/tests_$R.log
timeout 150 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -2 gpurun_out/bench_$R.err; cut -c1-400 gpurun_out/bench_$R.json
EMDEE_DEBUG=1 timeout 120 python bench.py --workload c4 --no-cpu-baseline --e2e-iters 1 > gpurun_out/bench_c4_$R.json 2> gpurun_out/bench_c4_$R.err; grep -v "^\[emdee\] force\|slab" gpurun_out/bench_c4_$R.err | tail -8; cut -c1-300 gpurun_out/bench_c4_$R.json
timeout 100 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-iters 1 > gpurun_out/bench_long_$R.json 2> gpurun_out/bench_long_$R.err; tail -2 gpurun_out/bench_long_$R.err; cut -c1-300 gpurun_out/bench_long_$R.json
