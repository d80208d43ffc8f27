# Example of an A/B sweep on the GPU box: per-kernel timings (EMDEE_DEBUG) and ms/step for a few settings.
cd /root/repo
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -3
B="timeout 300 python bench.py --no-cpu-baseline"
run() { # name, env..., extra args after --
  name=$1; shift
  env EMDEE_DEBUG=1 "$@" > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err
  echo "$name: $(grep 'bricks' gpurun_out/s_$name.err | sed 's/.emdee. //;s/(full.*//' | sort | uniq -c | tr '\n' ';') $(grep 'force kernel mode' gpurun_out/s_$name.err | head -3 | sed 's/.*mode//' | tr '\n' ';') $(python -c "
import json; d=json.loads(open('gpurun_out/s_$name.json').read().strip().splitlines()[-1]); print('ms/step %.3f e2e %.3f ms'%(d['ms_per_step'], d['e2e']['ms_per_call']))" 2>&1 | tail -1)"
}
run default $B
run skin40 $B --skin 0.4
