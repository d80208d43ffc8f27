# multi-GPU pass of the round (usage: bash tools/gpu_r2_8.sh <ranks>; SKIP60=1 leaves the 60-step line out): slab worker (parity against the oracle) at 8 ranks, config 3 and config 5 bench lines at 8 ranks
cd /root/repo
mkdir -p gpurun_out
W=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
SLAB_N=${SLAB_N:-40} SLAB_NDIV=1 timeout 280 $T --master-port 29541 tests/slab_worker.py > gpurun_out/slab_w$W.log 2>&1
grep -o "SLAB_RESULT.*" gpurun_out/slab_w$W.log | cut -c1-1500; tail -2 gpurun_out/slab_w$W.log | cut -c1-300
line() { python -c "
import json
d=json.loads([l for l in open('gpurun_out/$1.json') if l.startswith('{')][-1]); print('$1: value %.4g ms/step %.4f kernel %.4f ms build %.4f e2e ms/call %.2f bricks %s rebins %s parity %s'%(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['e2e']['ms_per_call'], d['run']['brick_cells'], d['run']['rebins_in_timed_steps'], (d['parity'] or {}).get('ok')))" 2>&1 | tail -1; }
EMDEE_DEBUG=2 timeout 200 $T --master-port 29511 bench.py --gpus $W --no-cpu-baseline > gpurun_out/scale_$W.json 2> gpurun_out/scale_$W.err; line scale_$W
grep "re-binning.*phase\|phases" gpurun_out/scale_$W.err | tail -6 | cut -c1-300
[ "$SKIP60" = 1 ] || timeout 200 $T --master-port 29513 bench.py --gpus $W --no-cpu-baseline --steps 60 --warmup 10 --e2e-iters 1 --no-parity > gpurun_out/scale_${W}_60.json 2> gpurun_out/scale_${W}_60.err; [ "$SKIP60" = 1 ] || line scale_${W}_60
timeout 280 $T --master-port 29512 bench.py --gpus $W --no-cpu-baseline --e2e-iters 1 --workload c5 --steps 10 --warmup 3 > gpurun_out/c5_$W.json 2> gpurun_out/c5_$W.err; line c5_$W
