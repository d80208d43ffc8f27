cd /root/repo
B="timeout 300 python bench.py --no-cpu-baseline --e2e-iters 1"
run() { # name, env..., extra args after --
  name=$1; shift
  env EMDEE_DEBUG=1 "$@" > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err
  echo "$name: $(grep 'bricks\|force kernel mode' gpurun_out/s_$name.err | head -3 | sed 's/.*mode//;s/.emdee. //;s/(full.*//' | tr '\n' ';') $(python -c "
import json; d=json.loads(open('gpurun_out/s_$name.json').read().strip().splitlines()[-1]); print('ms/step %.3f'%(d['ms_per_step']))" 2>&1 | tail -1)"
}
run a40 $B --skin 0.40 --rebin-every -1 --steps 60 --warmup 20
run a45 $B --skin 0.45 --rebin-every -1 --steps 60 --warmup 20
run a50 $B --skin 0.50 --rebin-every -1 --steps 60 --warmup 20
run a60 $B --skin 0.60 --rebin-every -1 --steps 60 --warmup 20
run a35 $B --skin 0.35 --rebin-every -1 --steps 60 --warmup 20
