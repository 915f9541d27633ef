#!/bin/bash
# Run a torchrun job under a hard watchdog that kills the whole process group; logs go to files.
#   tools/mgpu_run.sh <nproc> <seconds> <logfile> <script and args...>
N=$1; LIMIT=$2; LOG=$3; shift 3
mkdir -p "$(dirname "$LOG")"
setsid python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port ${PORT:-29533} "$@" > "$LOG" 2>&1 &
PID=$!
for ((i = 0; i < LIMIT; i++)); do
  if ! kill -0 $PID 2>/dev/null; then wait $PID; echo "exit code $?"; exit 0; fi
  sleep 1
done
echo "WATCHDOG: killing process group $PID after ${LIMIT}s"
kill -KILL -- -$PID 2>/dev/null
sleep 1
exit 124
