# compute-sanitizer pass over the small-N GPU tests (the reference has no race / bounds checking at all).
# Not run in round 1 (the GPU budget went to parity, bench and ncu); first validation step of the next round:
#   gpurun --timeout 900 -- 'bash tools/gpu_sanitize.sh r2'
# memcheck: out-of-bounds / misaligned accesses (tail lanes, staging past the brick capacity, list chunks);
# racecheck: shared-memory hazards (the producer/consumer hand-over of k_force_list_p uses named barriers, the list build's
#            rows are written with inline-asm stores the compiler does not see); synccheck: barrier misuse.
cd /root/repo
R=${1:-r2}
SEL="cutoff_fixture or cells_bit_exact or allpairs_ragged or mixed_lj or velocity_verlet or pair_list_shell or exclusions_molecular or tiny_box"
for tool in memcheck racecheck synccheck; do
  timeout 800 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitize_${tool}_$R.log 2>&1
  echo "$tool: exit $? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_$R.log) summaries; $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_$R.log | sort | uniq -c | tail -3)"
done
