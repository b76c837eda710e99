#!/bin/bash
# usage: gpu_retry.sh <logfile> <timeout> <ngpus> <cmd>   (retries while the pod answers "transient / busy")
log=$1; shift; to=$1; shift; n=$1; shift
for i in $(seq 1 40); do
  if [ "$n" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  else /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1; fi
  if ! grep -q "status=transient\|status=busy\|no box\|rc=3" $log; then exit 0; fi
  sleep 60
done
