#!/bin/bash
# usage: gpurun_retry.sh <log> <gpus> <timeout_s> <command...>  -- retries while the pod answers "busy" (nothing is charged then)
log=$1; gpus=$2; to=$3; shift 3
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 120; continue; fi
  break
done
echo "tries: $try" >> $log
