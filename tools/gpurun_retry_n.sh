#!/usr/bin/env bash
# tools/gpurun_retry_n.sh GPUS TIMEOUT_S COMMAND -- like gpurun_retry.sh, on GPUS GPUs of one box
g=$1; t=$2; shift 2
for attempt in $(seq 1 20); do
    /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 120
done
exit 3
