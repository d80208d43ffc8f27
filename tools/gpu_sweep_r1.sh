cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-iters 1"
for cfg in "4,2,2 256" "4,2,2 192" "4,4,2 256"; do
  set -- $cfg
  EMDEE_DEBUG=1 EMDEE_BRICK=$1 EMDEE_LBLOCK=$2 $B > gpurun_out/s_$1_$2.json 2> gpurun_out/s_$1_$2.err
  echo "$cfg:"; grep "force kernel mode" gpurun_out/s_$1_$2.err; python -c "
import json; d=json.loads(open('gpurun_out/s_$1_$2.json').read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'])"
done
EMDEE_BRICK=4,2,2 EMDEE_LBLOCK=256 ncu --set full --clock-control none --import-source on -k regex:k_list_build -s 1 -c 1 -o gpurun_out/prof_r1_build1 $B > gpurun_out/ncu_build1.log 2>&1; tail -2 gpurun_out/ncu_build1.log
EMDEE_BRICK=4,2,2 EMDEE_LBLOCK=256 ncu --set full --clock-control none --import-source on -k regex:k_force_list -s 8 -c 1 -o gpurun_out/prof_r1_list3 $B > gpurun_out/ncu_list3.log 2>&1; tail -2 gpurun_out/ncu_list3.log
