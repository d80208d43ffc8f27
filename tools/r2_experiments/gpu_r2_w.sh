# emdee_compute_nonbonded_into (chunked single-point evaluation, results copied out behind the compute): tests and the e2e leg
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
B="timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-iters 5"
for p in 8 0 4 16; do
EMDEE_PIPE=$p $B > gpurun_out/w_pipe$p.json 2> gpurun_out/w_pipe$p.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/w_pipe$p.json') if l.startswith('{')][-1]); print('EMDEE_PIPE=$p: ms/step %.4f kernel %.4f e2e %.3f ms/call parity %s'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['ms_per_call'], d['parity']['ok'] if d.get('parity') else None))" 2>&1 | tail -1
done
