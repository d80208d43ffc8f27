# Example of an A/B sweep on the GPU box: per-kernel timings (EMDEE_DEBUG) and ms/step for a few settings.
cd /root/repo
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -4
B="timeout 300 python bench.py --no-cpu-baseline --e2e-iters 1"
run() { # name, env..., extra args after --
  name=$1; shift
  env EMDEE_DEBUG=1 "$@" > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err
  echo "$name: $(grep 'force kernel mode' gpurun_out/s_$name.err | head -3 | sed 's/.*mode//' | tr '\n' ';') $(python -c "
import json; d=json.loads(open('gpurun_out/s_$name.json').read().strip().splitlines()[-1]); print('ms/step %.3f launches %d'%(d['ms_per_step'], d['gpu_launches']))" 2>&1 | tail -1)"
}
run fusevv1 EMDEE_FUSE_VV=1 $B
run fusevv0 EMDEE_FUSE_VV=0 $B
