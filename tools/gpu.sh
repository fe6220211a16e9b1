#!/bin/bash
# tools/gpu.sh [--gpus N] [--timeout S] -- 'command'   -- gpurun with retries while the pod answers "busy"/"backing off"
# (exit 3 = nothing charged).  Log of the last attempt: gpurun_out/<name>.log via caller redirection.
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
