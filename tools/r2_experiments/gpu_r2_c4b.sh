# split lists (two lanes per home atom) on top of compacted staging: the whole GPU suite, config 4 with / without the split, config 3 check
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for sp in 1 0; do
EMDEE_DEBUG=1 EMDEE_SPLIT=$sp timeout 200 python bench.py --workload c4 --no-cpu-baseline --e2e-iters 1 > gpurun_out/c4_split$sp.json 2> gpurun_out/c4_split$sp.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/c4_split$sp.json') if l.startswith('{')][-1]); r=d['roofline']; print('EMDEE_SPLIT=$sp: ms/step %.4f %s %.4f ms frac %.4f build %.4f e2e %.2f parity %s'%(d['ms_per_step'], r['kernel'], r['ms_per_launch'], r['frac'], r['list_build']['ms_per_launch'], d['e2e']['ms_per_call'], (d.get('parity') or {}).get('ok')))" 2>&1 | tail -1
done
timeout 150 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-iters 1 > gpurun_out/c3_after_dense.json 2> gpurun_out/c3_after_dense.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/c3_after_dense.json') if l.startswith('{')][-1]); r=d['roofline']; print('c3: ms/step %.4f kernel %.4f frac %.4f parity %s'%(d['ms_per_step'], r['ms_per_launch'], r['frac'], d['parity']['ok']))" 2>&1 | tail -1
