cd /root/repo
timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -8
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s_e2e.json 2> gpurun_out/s_e2e.err; tail -3 gpurun_out/s_e2e.err
python -c "
import json; d=json.loads(open('gpurun_out/s_e2e.json').read().strip().splitlines()[-1]); print('ms/step %.3f e2e %.4g pairs/s, %.3f ms/call'%(d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_call']))"
