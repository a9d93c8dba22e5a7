#!/bin/bash
# usage: scripts/gpurun_retry.sh <tries> <gpurun args...>   -- retries while gpurun answers "no box / slot free" (exit 3)
tries=$1; shift
for i in $(seq 1 "$tries"); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpurun_retry] attempt $i: busy, sleeping 60 s" >&2
    sleep 60
done
exit 3
