#!/bin/bash
# gpurun with retries on "no box / slot free right now" (exit code 3, nothing charged).
# Usage: tools/gpurun_retry.sh [gpurun options] -- '<command>'
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt answered busy; sleeping 150 s" >&2
  sleep 150
done
exit 3
