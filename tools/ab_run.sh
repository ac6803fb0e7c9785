#!/bin/bash
# On the GPU box: times every variant built by tools/ab_build.sh (and the default library) on the given tools/perf.py cases,
# then runs the GPU parity suite on each variant named after "--test".
#   gpurun -- 'bash tools/ab_run.sh "28 7 msb32,lsb32v4" default segconst --test segconst'
set -u
cases="$1"; shift
mkdir -p gpurun_out; out=gpurun_out/ab_$(date +%H%M%S).jsonl
testing=0
for v in "$@"; do
  if [ "$v" = "--test" ]; then testing=1; continue; fi
  lib=""; [ "$v" != "default" ] && lib="gpu_sort_b200/variants/$v.so"
  if [ $testing = 0 ]; then
    echo "# $v" >> $out; B200SORT_LIB=$lib python tools/perf.py $cases >> $out 2>&1
  else
    echo "# pytest $v" >> $out; B200SORT_LIB=$lib timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
  fi
done
cut -c1-160 $out
