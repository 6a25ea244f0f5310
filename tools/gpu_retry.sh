#!/bin/bash
# gpurun with retry while the pod is busy (exit code 3).  usage: tools/gpu_retry.sh <timeout_s> '<command>' [extra gpurun flags]
t=$1; shift; cmd=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@" -- "$cmd"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
