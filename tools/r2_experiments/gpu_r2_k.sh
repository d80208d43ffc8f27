cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/tests_r2k.log; tail -6 gpurun_out/tests_r2k.log
B="timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 1"
$B > gpurun_out/n3_off.json 2> gpurun_out/n3_off.err
EMDEE_N3=1 $B > gpurun_out/n3_on.json 2> gpurun_out/n3_on.err; tail -2 gpurun_out/n3_on.err
for v in off on; do python -c "
import json
d=json.loads([l for l in open('gpurun_out/n3_$v.json') if l.startswith('{')][-1]); print('N3 $v: ms/step %.4f kernel %.4f build %.4f parity %s frac %.4f'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['parity']['ok'] if d['parity'] else None, d['roofline']['frac']))"; done
