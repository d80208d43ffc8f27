# Round 2, first GPU call (one GPU): the full GPU test suite, compute-sanitizer over the small-N tests, the default bench
# (now with the `parity` block), BASELINE config 2 as stated (1000 steps at N = 256k), config 5 on one GPU, one A/B variant.
cd /root/repo
R=${1:-r2a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/tests_$R.log; tail -8 gpurun_out/tests_$R.log
timeout 200 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -2 gpurun_out/bench_$R.err; cut -c1-600 gpurun_out/bench_$R.json
timeout 200 python bench.py --workload c2 --steps 1000 --warmup 100 > gpurun_out/bench_c2_$R.json 2> gpurun_out/bench_c2_$R.err; tail -2 gpurun_out/bench_c2_$R.err; cut -c1-400 gpurun_out/bench_c2_$R.json
EMDEE_DEBUG=1 timeout 400 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --e2e-iters 1 > gpurun_out/bench_c5_$R.json 2> gpurun_out/bench_c5_$R.err; grep -v "force kernel mode" gpurun_out/bench_c5_$R.err | tail -4; cut -c1-400 gpurun_out/bench_c5_$R.json
bash tools/gpu_variants.sh preload2 2>&1 | tail -3
SEL="cutoff_fixture or cells_bit_exact or allpairs_ragged or mixed_lj or velocity_verlet or pair_list_shell or exclusions_molecular or tiny_box or state_invalidation"
for tool in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitize_${tool}_$R.log 2>&1
  echo "$tool: exit $? ; $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_$R.log | sort | uniq -c | tail -3); $(tail -1 gpurun_out/sanitize_${tool}_$R.log)"
done
