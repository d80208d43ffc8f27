# compacted staging for every persistent configuration on one GPU (EMDEE_COMPACT=2) against the default (dense cells only):
# the whole GPU suite with it, then config 3 both ways
cd /root/repo
EMDEE_COMPACT=2 timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
B="timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 3"
for cpt in 2 1; do
EMDEE_DEBUG=1 EMDEE_COMPACT=$cpt $B > gpurun_out/c3_compact$cpt.json 2> gpurun_out/c3_compact$cpt.err
grep "bricks" gpurun_out/c3_compact$cpt.err | head -1 | cut -c1-250
python -c "
import json
d=json.loads([l for l in open('gpurun_out/c3_compact$cpt.json') if l.startswith('{')][-1]); r=d['roofline']; print('EMDEE_COMPACT=$cpt: ms/step %.4f kernel %.4f frac %.4f build %.4f e2e %.3f parity %s'%(d['ms_per_step'], r['ms_per_launch'], r['frac'], r['list_build']['ms_per_launch'], d['e2e']['ms_per_call'], d['parity']['ok']))" 2>&1 | tail -1
done
