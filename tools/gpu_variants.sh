# A/B of compile-time tuning variants (build/libemdee_<name>.so built with -D...; selected with EMDEE_B200_LIB).
cd /root/repo
B="timeout 300 python bench.py --no-cpu-baseline --e2e-iters 1"
run() { # name, env..., extra args after --
  name=$1; shift
  env EMDEE_DEBUG=1 "$@" > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err
  echo "$name: $(grep 'bricks' gpurun_out/s_$name.err | head -1 | sed 's/.emdee. //;s/(full.*//') $(grep 'force kernel mode' gpurun_out/s_$name.err | head -3 | sed 's/.*mode//' | tr '\n' ';') $(python -c "
import json; d=json.loads(open('gpurun_out/s_$name.json').read().strip().splitlines()[-1]); print('ms/step %.3f'%(d['ms_per_step']))" 2>&1 | tail -1)"
}
run base $B
for v in "$@"; do run $v EMDEE_B200_LIB=/root/repo/build/libemdee_$v.so $B; done
