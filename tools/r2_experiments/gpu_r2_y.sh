# two-level list (prune / replay): GPU tests, then the bench with skin2 = 0 (off), default, and two other values
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
B="timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 1"
line() { python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/$1.json') if l.startswith('{')][-1]); print('$1: ms/step %.4f kernel %.4f build %.4f parity %s frac %.4f modes %s'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['parity']['ok'] if d.get('parity') else None, d['roofline']['frac'], d['roofline']['launches_by_list_mode']))" 2>&1 | tail -1; }
for k in 0 0.12 0.08 0.18; do
EMDEE_SKIN2=$k $B > gpurun_out/y_skin2_$k.json 2> gpurun_out/y_skin2_$k.err; line y_skin2_$k; tail -2 gpurun_out/y_skin2_$k.err | cut -c1-300
done
