#!/bin/bash
# Builds kernel variants for an A/B run on the GPU box (they travel with the gpurun snapshot; gpu_sort_b200/variants/ is git-ignored).
#   tools/ab_build.sh name1:"-DFLAG=1" name2:"-DOTHER=8 -DMORE=1" ...
# Each variant becomes gpu_sort_b200/variants/<name>.so; select it with B200SORT_LIB=gpu_sort_b200/variants/<name>.so
# (gpu_sort_b200/__init__.py honours it for tests, tools/perf.py and bench.py alike).
# Compile-time switches that exist today: B200_SEG_CONST (experimental), B200_HIST_GROUP, B200_HIST_TICKET,
# B200_PDL, B200_LOCAL_LEAN_EXACT, B200_LOCAL_VEC_OUT, B200_LOCAL_OCC384, B200_LOCAL_IPT32, B200_LOCAL_THREADS32,
# B200_SCATTER_THREADS, B200_SCATTER_OCC, B200_IPT_NUM / B200_IPT_DEN, B200SORT_HW_MATCH.
set -e
cd "$(dirname "$0")/../gpu_sort_b200/csrc"
mkdir -p ../variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  echo "== $name: $flags"
  make -j"$(nproc)" OUT=../variants/$name.so OBJDIR=build_$name EXTRA="$flags" 2>&1 | grep -E "error|Error" || true
  ls -la ../variants/$name.so
done
