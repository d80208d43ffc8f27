# claim-at-end fix of the mbarrier hand-over: A/B against named barriers, the cross-brick carry, 3 producers, role timers
cd /root/repo
mkdir -p gpurun_out
B="timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 1"
line() { python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/$1.json') if l.startswith('{')][-1]); print('$1: ms/step %.4f kernel %.4f build %.4f parity %s frac %.4f'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['parity']['ok'] if d.get('parity') else None, d['roofline']['frac']))" 2>&1 | tail -1; }
$B > gpurun_out/u_mbar1.json 2> gpurun_out/u_mbar1.err; line u_mbar1
for v in mbar0 carry np3c; do
EMDEE_B200_LIB=/root/repo/build/libemdee_$v.so $B > gpurun_out/u_$v.json 2> gpurun_out/u_$v.err; line u_$v
done
for t in timing timingc; do
EMDEE_DEBUG=1 EMDEE_B200_LIB=/root/repo/build/libemdee_$t.so $B --no-parity > gpurun_out/u_$t.json 2> gpurun_out/u_$t.err; echo "$t: $(grep 'role timers' gpurun_out/u_$t.err | tail -1)"
done
EMDEE_B200_LIB=/root/repo/build/libemdee_carry.so timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
