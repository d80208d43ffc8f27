cd /root/repo
mkdir -p gpurun_out
EMDEE_DEBUG=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab" --timeout 300 2>&1 | tail -3
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --e2e-iters 1"
EMDEE_DEBUG=2 timeout 200 $T --no-parity > gpurun_out/p2_dbg2.json 2> gpurun_out/p2_dbg2.err; grep "slab re-binnings" gpurun_out/p2_dbg2.err | tail -4
EMDEE_DEBUG=1 timeout 200 $T --steps 60 > gpurun_out/p2_60.json 2> gpurun_out/p2_60.err; grep "force kernel mode" gpurun_out/p2_60.err | tail -2; python -c "
import json
d=json.loads([l for l in open('gpurun_out/p2_60.json') if l.startswith('{')][-1]); print('N=2 60 steps: ms/step %.4f kernel %.4f build %.4f rebins %s parity %s'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['run']['rebins_in_timed_steps'], d['parity']['ok']))"
for v in cur r1 cur r1; do
L=/root/repo/emdee.jl_b200/csrc/libemdee_b200.so; [ $v = r1 ] && L=/root/repo/build/libemdee_r1.so
EMDEE_B200_LIB=$L timeout 200 python bench.py --no-cpu-baseline --no-parity --e2e-iters 1 --steps 60 > gpurun_out/p1_$v.json 2> gpurun_out/p1_$v.err;  python -c "
import json
d=json.loads([l for l in open('gpurun_out/p1_$v.json') if l.startswith('{')][-1]); print('N=1 $v 60 steps: ms/step %.4f kernel %.4f build %.4f rebins %s'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['run'].get('rebins_in_timed_steps') if 'run' in d else d['config'].get('rebins_in_timed_steps')))"
done
