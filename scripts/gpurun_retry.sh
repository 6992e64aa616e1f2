#!/bin/bash
# scripts/gpurun_retry.sh LOGFILE TIMEOUT [--gpus N] -- 'command': retries a gpurun call while the pool answers "busy" (exit 3 / transient)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log" && [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
