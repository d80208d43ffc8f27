cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
B="timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 3"
EMDEE_DEBUG=1 EMDEE_TMA=1 $B > gpurun_out/tma_on.json 2> gpurun_out/tma_on.err; grep "TMA staging" gpurun_out/tma_on.err | tail -1; tail -1 gpurun_out/tma_on.err | cut -c1-300
EMDEE_TMA=0 $B > gpurun_out/tma_off.json 2> gpurun_out/tma_off.err
for v in off on; do python -c "
import json
d=json.loads([l for l in open('gpurun_out/tma_$v.json') if l.startswith('{')][-1]); print('TMA $v: ms/step %.4f kernel %.4f build %.4f parity %s frac %.4f e2e %.2f ms'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['parity']['ok'] if d['parity'] else None, d['roofline']['frac'], d['e2e']['ms_per_call']))"; done
