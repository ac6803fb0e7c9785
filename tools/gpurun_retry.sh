#!/bin/bash
# Retries a gpurun call while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh <timeout> '<command>' [gpus]
t="$1"; cmd="$2"; g="${3:-1}"
for i in $(seq 1 40); do
  if [ "$g" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$t" -- "$cmd"; else /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$cmd"; fi
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
