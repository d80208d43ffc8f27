cd /root/repo
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --e2e-iters 1"
EMDEE_DEBUG=2 timeout 200 $T --no-parity > gpurun_out/p2_dbg2.json 2> gpurun_out/p2_dbg2.err; grep "re-binning [0-9]* phase" gpurun_out/p2_dbg2.err | tail -15
EMDEE_DEBUG=1 timeout 200 $T --steps 60 > gpurun_out/p2_60.json 2> gpurun_out/p2_60.err; grep "force kernel mode" gpurun_out/p2_60.err | tail -2; python -c "
import json
d=json.loads([l for l in open('gpurun_out/p2_60.json') if l.startswith('{')][-1]); print('N=2 60 steps: ms/step %.4f kernel %.4f build %.4f rebins %s parity %s'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['run']['rebins_in_timed_steps'], d['parity']['ok']))"
timeout 200 python bench.py --no-cpu-baseline --no-parity --e2e-iters 1 --steps 60 > gpurun_out/p1_cur.json 2> gpurun_out/p1_cur.err;  python -c "
import json
d=json.loads([l for l in open('gpurun_out/p1_cur.json') if l.startswith('{')][-1]); print('N=1 60 steps: ms/step %.4f kernel %.4f build %.4f'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch']))"
