# Multi-GPU pass: slab-decomposition parity test and the scaling bench at N = $1
cd /root/repo
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab" 2>&1 | tail -5
for n in 1 $N; do
  if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  tail -2 gpurun_out/scale_$n.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_$n.json') if l.startswith('{')][-1]); print('N=$n value %.4g ms/step %.4f kernel %.4f ms e2e %.4g'%(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value']))"
done
