cd /root/repo
mkdir -p gpurun_out
run() { # world ndiv n
  SLAB_N=$3 SLAB_NDIV=$2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29541 tests/slab_worker.py > gpurun_out/w_$1_$2_$3.log 2>&1
  echo "world $1 ndiv $2 n $3: rc $? $(grep -o 'SLAB_RESULT.*' gpurun_out/w_$1_$2_$3.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l[12:]); print({k: d[k] for k in ('ok','full_force_err','window_force_err','window_E_err','window','window_bad_rows','id_window_rows_sum','vv_force_err_same_positions')})")"
}
run 4 2 16
run 4 1 16
run 2 2 16
run 4 1 24
