#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3, nothing charged):  tools/gpu_retry.sh <timeout_s> '<command>' [gpus]
t=$1; cmd=$2; gpus=${3:-1}
for i in $(seq 1 12); do
    if [ "$gpus" = 1 ]; then /usr/local/graft/bin/gpurun --timeout "$t" -- "$cmd"; else /usr/local/graft/bin/gpurun --gpus "$gpus" --timeout "$t" -- "$cmd"; fi
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 90
done
exit 3
