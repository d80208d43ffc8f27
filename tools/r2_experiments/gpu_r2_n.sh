# 8-GPU pass: slab parity test at 2 / 4 / 8 ranks, config 3 and config 5 bench lines at 8 ranks
cd /root/repo
mkdir -p gpurun_out
bash tools/gpu_multi.sh 8
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-cpu-baseline --e2e-iters 1"
EMDEE_DEBUG=1 timeout 280 $T --workload c5 --steps 10 --warmup 3 > gpurun_out/c5_8.json 2> gpurun_out/c5_8.err
grep "slab re-binning\|bricks" gpurun_out/c5_8.err | tail -3; python -c "
import json
d=json.loads([l for l in open('gpurun_out/c5_8.json') if l.startswith('{')][-1]); print('c5 N=8 value %.4g ms/step %.4f kernel %.4f ms build %.4f e2e ms/call %.2f parity %s'%(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['e2e']['ms_per_call'], d['parity']))" 2>&1 | tail -1
