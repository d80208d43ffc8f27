# mbarrier hand-over (FLP_MBAR=1, the new default) against the named-barrier hand-over: GPU tests, bench A/B, role timers
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
B="timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 1"
line() { python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/$1.json') if l.startswith('{')][-1]); print('$1: ms/step %.4f kernel %.4f build %.4f parity %s frac %.4f'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['parity']['ok'] if d.get('parity') else None, d['roofline']['frac']))" 2>&1 | tail -1; }
$B > gpurun_out/u_mbar1.json 2> gpurun_out/u_mbar1.err; line u_mbar1
EMDEE_B200_LIB=/root/repo/build/libemdee_mbar0.so $B > gpurun_out/u_mbar0.json 2> gpurun_out/u_mbar0.err; line u_mbar0
EMDEE_B200_LIB=/root/repo/build/libemdee_np3.so $B --no-parity > gpurun_out/u_np3.json 2> gpurun_out/u_np3.err; line u_np3
for t in timing timing0; do
EMDEE_DEBUG=1 EMDEE_B200_LIB=/root/repo/build/libemdee_$t.so $B --no-parity > gpurun_out/u_$t.json 2> gpurun_out/u_$t.err; echo "$t: $(grep 'role timers' gpurun_out/u_$t.err | tail -1)"
done
