#!/bin/bash
# Builds the reference's OWN drivers, source files untouched, against libb200sort.so through the header shims
# (INTEGRATION.md): lsb/sort.cu (cub::DeviceRadixSort call shape) and msb/src/test.cu (rdxsrt_unstable_sort).
# Needs /root/reference (build container only); the binaries land in oracle/_ref/ and travel to the GPU box.
set -e
cd "$(dirname "$0")/.."
REF=${REF:-/root/reference}
OUT=oracle/_ref
mkdir -p $OUT
FLAGS="-std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -w -I include/shim -I include -L gpu_sort_b200 -lb200sort -lcurand -Xlinker -rpath=\$ORIGIN/../../gpu_sort_b200"
nvcc $FLAGS -I $REF/lsb -I $REF/lsb/cub/test $REF/lsb/sort.cu -o $OUT/lsb_sort_on_b200sort
nvcc $FLAGS -include include/shim/sort/gpu_radix_sort.h -I $REF/msb/src $REF/msb/src/test.cu -o $OUT/msb_test_on_b200sort
ls -la $OUT/lsb_sort_on_b200sort $OUT/msb_test_on_b200sort
if [ "${1:-}" = "gtest" ]; then
  # the reference's own gtest suite (msb/tests/*.cu: 12 entropy levels x key/value types x sizes, oracle = cub::DeviceRadixSort)
  GT=$REF/msb/submodules/googletest/googletest
  g++ -O2 -c -I $GT/include -I $GT $GT/src/gtest-all.cc -o /tmp/gtest-all.o
  nvcc $FLAGS -include include/shim/sort/gpu_radix_sort.h -I $REF/msb/src -I $REF/msb/external -I $REF/msb/tests -I $GT/include \
    $REF/msb/tests/main.cu $REF/msb/tests/test_sort_keys.cu $REF/msb/tests/test_sort_pairs.cu \
    $REF/msb/external/benchmark/benchmark.cu $REF/msb/external/benchmark/get_real_time.cu /tmp/gtest-all.o \
    -lpthread -o $OUT/msb_gtests_on_b200sort
  ls -la $OUT/msb_gtests_on_b200sort
fi
