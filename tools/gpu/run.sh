#!/bin/bash
# usage: tools/gpu/run.sh [--gpus N] TIMEOUT 'command'   -- retries while the pod answers "no slot right now" (exit code 3)
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" > /tmp/gpurun_last.out 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.out; then break; fi
  sleep 90
done
cat /tmp/gpurun_last.out
exit $rc
