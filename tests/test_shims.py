"""CPU test of the C++ drop-in surface: a translation unit written against the REFERENCE's names -- cub::DeviceRadixSort
(lsb/cub/cub/device/device_radix_sort.cuh), cub::DeviceSegmentedRadixSort (device_segmented_radix_sort.cuh) and
rdxsrt_unstable_sort / RDXSRT_SortedSequence (msb/src/sort/gpu_radix_sort.h:169-197) -- compiles against the header shims
(include/shim) and links against libb200sort.so.  Only the size queries run (host arithmetic: no GPU is needed); the sorts
themselves are exercised on the GPU box by the reference's own drivers and gtest suite built the same way
(tools/build_ref_on_b200.sh, profiles/r01_ref_*_on_b200sort.log) and by tests/test_gpu_parity.py."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpu_sort_b200", "libb200sort.so")

SRC = r'''
#include <cstdio>
#include <cub/util_type.cuh>
#include <cub/device/device_radix_sort.cuh>                 // -> include/shim/cub/device/device_radix_sort.cuh
#include <cub/device/device_segmented_radix_sort.cuh>       // -> include/shim/cub/device/device_segmented_radix_sort.cuh
#include <shim/sort/gpu_radix_sort.h>                       // stands in for msb/src/sort/gpu_radix_sort.h

template <typename K, typename V> static size_t lsb_pairs_bytes(int n) {
  cub::DoubleBuffer<K> k; cub::DoubleBuffer<V> v; size_t b = 0;
  if (cub::DeviceRadixSort::SortPairs(nullptr, b, k, v, n) != cudaSuccess) return 0;
  size_t d = 0;
  if (cub::DeviceRadixSort::SortPairsDescending(nullptr, d, (const K*)nullptr, (K*)nullptr, (const V*)nullptr, (V*)nullptr, n, 3, 17) != cudaSuccess) return 0;
  return d >= b ? b : 0;                  // the pointer overloads carry a third buffer (dispatch_radix_sort.cuh:1099-1104)
}
template <typename K> static size_t seg_keys_bytes(int n, int segs) {
  cub::DoubleBuffer<K> k; size_t b = 0; const int* off = nullptr;
  if (cub::DeviceSegmentedRadixSort::SortKeys(nullptr, b, k, n, segs, off, off + 1) != cudaSuccess) return 0;
  size_t d = 0;
  if (cub::DeviceSegmentedRadixSort::SortKeysDescending(nullptr, d, (const K*)nullptr, (K*)nullptr, n, segs, off, off + 1) != cudaSuccess) return 0;
  return d >= b ? b : 0;
}
int main() {
  // the MSB entry point and its result type must exist with the reference's template signature (never called here: it would sort)
  RDXSRT_SortedSequence<unsigned int, cub::NullType> (*f)(unsigned int*, cub::NullType*, unsigned int, unsigned int*, cub::NullType*,
      LocalSortConfigSet<unsigned int, cub::NullType>*, RDXSRT_GPUDataManager<unsigned int, cub::NullType, unsigned int, 0, 0, 8>*, cudaStream_t) =
      &rdxsrt_unstable_sort<unsigned int, cub::NullType, unsigned int>;
  RDXSRT_SortedSequence<unsigned long long, unsigned int> (*g)(unsigned long long*, unsigned int*, unsigned int, unsigned long long*, unsigned int*,
      LocalSortConfigSet<unsigned long long, unsigned int>*, RDXSRT_GPUDataManager<unsigned long long, unsigned int, unsigned int, 0, 0, 8>*, cudaStream_t) =
      &rdxsrt_unstable_sort<unsigned long long, unsigned int, unsigned int>;
  void (*h)(float*, unsigned long long, float*) = &rdxsrt_unstable_sort_keys<float>;
  printf("%zu %zu %zu %zu %d\n", lsb_pairs_bytes<float, unsigned int>(1 << 20), lsb_pairs_bytes<unsigned long long, unsigned long long>(1 << 20),
         seg_keys_bytes<double>(1 << 20, 1000), seg_keys_bytes<int>(0, 0), (f != nullptr) + (g != nullptr) + (h != nullptr));
  return 0;
}
'''


@pytest.mark.skipif(shutil.which("nvcc") is None or not os.path.exists(LIB), reason="needs nvcc and the built library")
def test_reference_call_shapes_compile_and_link_against_the_shims(tmp_path):
    src = tmp_path / "shim_tu.cu"
    src.write_text(SRC)
    exe = tmp_path / "shim_tu"
    cmd = ["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-I", os.path.join(ROOT, "include", "shim"),
           "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", os.path.join(ROOT, "gpu_sort_b200"), "-lb200sort",
           "-Xlinker", "-rpath=" + os.path.join(ROOT, "gpu_sort_b200")]
    subprocess.check_call(cmd)
    out = subprocess.check_output([str(exe)], text=True).split()
    a, b, c, d, n = (int(x) for x in out)
    assert a >= 256 and b > a and c >= 256 and d >= 256          # non-trivial sizes; wider keys and values need more
    assert n == 3
