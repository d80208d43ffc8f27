# role timers of the prune and the replay launches (timing build; EMDEE_DEBUG_LM forces the mode: timing only, the forces of a stale replay are wrong)
cd /root/repo
B="timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 1 --no-parity"
for lm in 1 2; do
EMDEE_DEBUG=1 EMDEE_DEBUG_LM=$lm EMDEE_B200_LIB=/root/repo/build/libemdee_timing.so $B > gpurun_out/z_lm$lm.json 2> gpurun_out/z_lm$lm.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/z_lm$lm.json') if l.startswith('{')][-1]); print('forced LM=$lm: ms/step %.4f kernel %.4f modes %s'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['launches_by_list_mode']))" 2>&1 | tail -1
grep 'role timers' gpurun_out/z_lm$lm.err | tail -1
done
