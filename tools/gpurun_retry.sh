#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> '<command>'   — retries while the pod answers "transient/busy" (exit 3)
log=$1; to=$2; cmd=$3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$cmd" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient\|nothing was charged" "$log"; then exit $rc; fi
  sleep 90
done
