#!/bin/bash
# One gpurun call with a retry ONLY when nothing ran (status transient / busy).  Never re-submit after a failure or a time-out: a command
# that hit its limit once will hit it again, and every attempt is charged -- a loop that retried on "not ok" cost this project 76 GPU-minutes
# of one round (DESIGN 4.4 d).  Wrap anything that might hang in its own `timeout` inside the command as well.
#   tools/gpu_once.sh TIMEOUT_SECONDS "command"
T=$1; CMD=$2
for i in 1 2 3 4 5 6; do
  out=$(timeout $((T+2400)) gpurun --timeout $T -- "$CMD" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 150; continue; fi
  exit 0
done
