#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout-seconds> [--gpus N] -- <command>
# retries gpurun while the pod answers "busy / no slot" (exit 3), every 90 s for up to ~40 min
log=$1; to=$2; shift 2
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 28); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "rc=$rc" >> "$log"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$log"
