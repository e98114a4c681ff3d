#!/bin/bash
# Development aid: runs a gpurun call, retrying while the pod answers "busy / draining" (exit code 3 or a transient status).
#   scripts/gpurun_retry.sh [--gpus N] TIMEOUT 'command'
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" 2>&1)
  RC=$?
  if echo "$OUT" | grep -q "status=transient\|status=busy" || [ $RC -eq 3 ]; then sleep 45; continue; fi
  echo "$OUT"; exit $RC
done
echo "$OUT"; exit 3
