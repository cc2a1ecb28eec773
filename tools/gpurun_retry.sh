#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <logfile> <command string>
# retries while the pod answers busy/transient (exit 3); dev helper, not part of the product
T=$1; LOG=$2; shift 2
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|no box\|retry in a few minutes" "$LOG" && [ $rc -ne 0 ]; then sleep 90; continue; fi
  exit $rc
done
exit 3
