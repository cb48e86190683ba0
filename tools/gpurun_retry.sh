#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod answers "transient" (nothing charged)
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then break; fi
  sleep 120
done
