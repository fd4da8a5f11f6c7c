#!/bin/bash
# gpurun with retry while the pod has no free slot (exit 3 / "transient"): nothing is charged for those.
# usage: tools/gpurun_retry.sh [--gpus N] --timeout S -- '<command>'
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > /tmp/gpurun_attempt.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_attempt.log || [ $rc -eq 3 ]; then
    echo "[retry] attempt $attempt: no slot, sleeping 90s"; sleep 90; continue
  fi
  cat /tmp/gpurun_attempt.log; exit $rc
done
cat /tmp/gpurun_attempt.log; exit 3
