#!/bin/bash
# usage: tools/gpu/submit.sh <log> <timeout_s> [--gpus N] -- <command>   : retries while the pod answers busy / transient
LOG=$1; shift; TMO=$1; shift
EXTRA=()
while [ "$1" != "--" ]; do EXTRA+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$TMO" "${EXTRA[@]}" -- "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$LOG" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "submit: finished rc=$rc after $i tries" >> "$LOG"
  exit $rc
done
echo "submit: gave up" >> "$LOG"
