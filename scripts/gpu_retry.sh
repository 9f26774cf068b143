#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> <out_file> <command...>   -- retries while the pod answers "busy/transient"
T=$1; OUT=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$OUT" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$OUT" || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
exit 3
