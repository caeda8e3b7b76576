#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   - retries while the pod answers "busy" (exit 3: nothing charged)
for attempt in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] busy (attempt $attempt), retrying in 120 s" >&2
  sleep 120
done
exit 3
