#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <timeout-seconds> [--gpus N] -- <command>
# Retries gpurun while the pod answers "busy / draining" (exit code 3, nothing charged).
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  gpurun --timeout "$to" "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 60
done
exit 3
