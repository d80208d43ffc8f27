cd /root/repo
EMDEE_DEBUG=1 timeout 400 python bench.py --workload c5 --no-cpu-baseline --steps 10 --warmup 3 --e2e-iters 1 > gpurun_out/s_c5.json 2> gpurun_out/s_c5.err; tail -4 gpurun_out/s_c5.err; python -c "
import json; d=json.loads(open('gpurun_out/s_c5.json').read().strip().splitlines()[-1]); print('c5 ms/step %.3f value %.4g e2e %.3f ms frac %.3f'%(d['ms_per_step'], d['value'], d['e2e']['ms_per_call'], d['roofline']['frac']))"
nvidia-smi --query-gpu=memory.used --format=csv
