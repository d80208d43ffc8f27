# after the mbarrier hand-over: brick shapes the lockstep model ruled out (more warp tasks than consumer warps), 640-thread blocks
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
B="timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 1 --no-parity"
line() { python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/$1.json') if l.startswith('{')][-1]); print('$1: ms/step %.4f kernel %.4f build %.4f frac %.4f'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['roofline']['frac']))" 2>&1 | tail -1; grep bricks gpurun_out/$1.err | tail -1; }
for sh in 4,3,2 6,2,2 3,3,2 5,2,2 3,2,2; do
EMDEE_DEBUG=1 EMDEE_BRICK=$sh $B > gpurun_out/v_$sh.json 2> gpurun_out/v_$sh.err; line v_$sh
done
for v in t640 t640b; do
EMDEE_DEBUG=1 EMDEE_B200_LIB=/root/repo/build/libemdee_$v.so $B > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err; line v_$v
done
