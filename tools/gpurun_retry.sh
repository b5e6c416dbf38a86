#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   retries while the pod answers "busy" (exit code 3: nothing charged)
log=$1; shift
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $try tries" >> "$log"; exit $rc; fi
  sleep 45
done
echo "gpurun still busy after 40 tries" >> "$log"; exit 3
