#!/usr/bin/env bash
# tools/gpurun_retry.sh TIMEOUT_S COMMAND -- run one gpurun call, retrying while the pod answers busy (exit 3: nothing charged).
t=$1; shift
for attempt in $(seq 1 20); do
    /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 120
done
exit 3
