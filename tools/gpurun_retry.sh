#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient): tools/gpurun_retry.sh [--gpus N] <timeout> '<command>'
GPUS=""
if [ "$1" == "--gpus" ]; then GPUS="--gpus $2"; shift; shift; fi
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $GPUS --timeout $T -- "$@" > /root/repo/gpurun_out_err.txt 2>&1
  if ! grep -q "status=transient" /root/repo/gpurun_out_err.txt; then break; fi
  sleep 150
done
tail -25 /root/repo/gpurun_out_err.txt
