# single-point e2e: list kernels (build + walk) against the window-scan kernel (EMDEE_LIST=0), plain sequence both
cd /root/repo
B="timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-iters 5 --no-parity"
for l in 1 0; do
EMDEE_DEBUG=1 EMDEE_PIPE=0 EMDEE_LIST=$l $B > gpurun_out/x_list$l.json 2> gpurun_out/x_list$l.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/x_list$l.json') if l.startswith('{')][-1]); print('EMDEE_LIST=$l: ms/step %.4f e2e %.3f ms/call'%(d['ms_per_step'], d['e2e']['ms_per_call']))" 2>&1 | tail -1
grep "force kernel mode" gpurun_out/x_list$l.err | tail -4
done
