# Multi-GPU pass: slab-decomposition parity test (unless SKIP_TEST=1) and the scaling bench at N = $1 [$2 ...]
cd /root/repo
mkdir -p gpurun_out
[ "$SKIP_TEST" = 1 ] || EMDEE_DEBUG=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "slab" --timeout 300 > gpurun_out/slab_test.log 2>&1
[ "$SKIP_TEST" = 1 ] || { grep -o "SLAB_RESULT.*" gpurun_out/slab_test.log | cut -c1-1800; grep "Error\|error:" gpurun_out/slab_test.log | head -5; tail -3 gpurun_out/slab_test.log; }
for n in "$@"; do
  if [ $n = 1 ]; then timeout 300 python bench.py --gpus 1 --no-cpu-baseline $BENCH_ARGS > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else EMDEE_DEBUG=1 timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu-baseline $BENCH_ARGS > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  grep "force kernel mode\|slab step\|peer-mapped\|status" gpurun_out/scale_$n.err | sort | uniq -c | head -12
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_$n.json') if l.startswith('{')][-1]); print('N=$n value %.4g ms/step %.4f kernel %.4f ms build %.4f e2e %.4g ms/call %.2f parity %s'%(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['e2e']['value'], d['e2e']['ms_per_call'], d['parity']))" 2>&1 | tail -1
done
