#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3): usage gpurun_retry.sh <timeout> [--gpus N] -- <command>
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
