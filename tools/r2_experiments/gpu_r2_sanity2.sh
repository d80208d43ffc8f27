# 2-GPU sanity of the final code: slab worker against the oracle, one short config-3 bench line with its parity block
cd /root/repo
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
SLAB_N=16 SLAB_NDIV=1 timeout 200 $T --master-port 29541 tests/slab_worker.py > gpurun_out/slab_w2_final.log 2>&1
grep -ao "SLAB_RESULT.*" gpurun_out/slab_w2_final.log | head -1 | cut -c1-400; tail -1 gpurun_out/slab_w2_final.log | cut -c1-200
timeout 150 $T --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --steps 20 --warmup 5 --e2e-iters 1 > gpurun_out/scale_2_final.json 2> gpurun_out/scale_2_final.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_2_final.json') if l.startswith('{')][-1]); print('2 GPUs: value %.4g ms/step %.4f kernel %.4f e2e %.2f parity %s'%(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['ms_per_call'], d['parity']['ok']))" 2>&1 | tail -1
