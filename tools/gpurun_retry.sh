#!/bin/bash
# usage: gpurun_retry.sh <log> <gpurun args...>: retries while the pod answers busy/transient (exit 3 or "transient")
log=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "finished rc=$rc" >> "$log"
