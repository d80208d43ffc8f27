# last GPU pass of the round: the whole GPU suite, config 3 (stepping + e2e) and config 4 on the final library
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
line() { python -c "
import json
d=json.loads([l for l in open('gpurun_out/$1.json') if l.startswith('{')][-1]); r=d['roofline']; print('$1: value %.4g ms/step %.4f %s %.4f ms frac %.4f build %.4f e2e %.3f parity %s'%(d['value'], d['ms_per_step'], r['kernel'], r['ms_per_launch'], r['frac'], r['list_build']['ms_per_launch'], d['e2e']['ms_per_call'], d['parity']['ok']))" 2>&1 | tail -1; }
timeout 150 python bench.py --no-cpu-baseline > gpurun_out/last_c3.json 2> gpurun_out/last_c3.err; line last_c3
timeout 200 python bench.py --workload c4 --no-cpu-baseline --e2e-iters 1 > gpurun_out/last_c4.json 2> gpurun_out/last_c4.err; line last_c4
