#!/bin/bash
# usage: tools/gpurun_retry.sh <tries> <gpurun args...> -- retries while gpurun answers "no box right now" (exit 3)
tries=$1; shift
for i in $(seq 1 "$tries"); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $i got exit 3; sleeping 150 s"
  sleep 150
done
exit 3
