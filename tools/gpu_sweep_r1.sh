cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-iters 1"
run() { # name, env..., extra args after --
  name=$1; shift
  env EMDEE_DEBUG=1 "$@" > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err
  echo "$name: $(grep 'force kernel mode' gpurun_out/s_$name.err | sed 's/.*mode//' | tr '\n' ';') $(python -c "
import json; d=json.loads(open('gpurun_out/s_$name.json').read().strip().splitlines()[-1]); print('ms/step %.3f'%d['ms_per_step'])" 2>&1 | tail -1)"
}
run b2_ilp8_192 EMDEE_BRICK=4,2,2 EMDEE_LBLOCK=192 EMDEE_ILP8=1 $B
run b2_s5r7 EMDEE_BRICK=4,2,2 EMDEE_LBLOCK=192 EMDEE_ILP8=1 $B --skin 0.5 --rebin-every 7 --steps 21
EMDEE_BRICK=4,2,2 EMDEE_LBLOCK=192 EMDEE_ILP8=1 ncu --set full --clock-control none --import-source on -k regex:k_list_build -s 1 -c 1 -o gpurun_out/prof_r1_build2 $B > gpurun_out/ncu_build2.log 2>&1; tail -1 gpurun_out/ncu_build2.log
