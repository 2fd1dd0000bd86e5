#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
# usage: tools/gpurun_retry.sh [--gpus N] <timeout_s> '<command>'
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout $T -- "$@"
  rc=$?
  # 3 = busy (nothing charged); 2 = refused, e.g. an abandoned earlier request still holds the one-call slot
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then exit $rc; fi
  sleep 90
done
exit 3
