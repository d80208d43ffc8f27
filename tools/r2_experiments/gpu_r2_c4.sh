# compacted staging (dense cells): the tests that touch it, then config 4 with and without it
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "dense_cells or molecular or config4 or pairs14 or state_invalidation or cutoff_fcc" 2>&1 | tail -6
for cpt in 1 0; do
EMDEE_DEBUG=1 EMDEE_COMPACT=$cpt timeout 200 python bench.py --workload c4 --no-cpu-baseline --e2e-iters 1 > gpurun_out/c4_compact$cpt.json 2> gpurun_out/c4_compact$cpt.err
grep "bricks" gpurun_out/c4_compact$cpt.err | head -1 | cut -c1-250
python -c "
import json
d=json.loads([l for l in open('gpurun_out/c4_compact$cpt.json') if l.startswith('{')][-1]); r=d['roofline']; print('EMDEE_COMPACT=$cpt: ms/step %.4f %s %.4f ms frac %.4f build %.4f e2e %.2f parity %s'%(d['ms_per_step'], r['kernel'], r['ms_per_launch'], r['frac'], r['list_build']['ms_per_launch'], d['e2e']['ms_per_call'], (d.get('parity') or {}).get('ok')))" 2>&1 | tail -1
done
