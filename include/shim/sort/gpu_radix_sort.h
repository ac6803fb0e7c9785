// include/shim/sort/gpu_radix_sort.h -- drop-in for the reference header msb/src/sort/gpu_radix_sort.h.
//
// Force-include it (`nvcc -include <repo>/include/shim/sort/gpu_radix_sort.h ...`): it claims the reference header's include
// guard, so the `#include "sort/gpu_radix_sort.h"` in msb/src/test.cu, msb/tests/test_sort_keys.cu and
// msb/tests/test_sort_pairs.cu expands to nothing and those files compile unchanged: the names they use -- rdxsrt_unstable_sort<KeyT,ValueT,IndexT>,
// RDXSRT_SortedSequence, rdxsrt_unstable_sort_keys / _pairs (gpu_radix_sort.h:169-184, 187-197, 510-587) -- are re-created
// here on top of the C ABI (include/b200sort.h, libb200sort.so).  Header-only, host code only.
#pragma once
#ifndef GPU_RADIX_SORT_H_
#define GPU_RADIX_SORT_H_          /* the reference header's guard (msb/src/sort/gpu_radix_sort.h:1-2) */
#endif
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <cub/util_type.cuh>          // cub::NullType, the reference's keys-only marker (any CUB provides it)
#include "../../b200sort.h"

// The reference header chain leaks these into every file that includes it (cuda_radix_sort.h:7-10 includes cub/cub.cuh and
// says `using namespace cub;`); its tests rely on that (msb/tests/test_sort_keys.cu:9,17).
#include <cub/cub.cuh>
using namespace cub;

template <typename KeyT, typename ValueT>
struct RDXSRT_SortedSequence {             // gpu_radix_sort.h:169-184
  KeyT* sorted_keys;
  ValueT* sorted_values;
};

// Opaque stand-ins so that call sites passing the optional trailing arguments still compile.
template <typename KeyT, typename ValueT> struct LocalSortConfigSet;
template <typename KeyT, typename ValueT, typename IndexT, unsigned TPB = 0, unsigned KPT = 0, unsigned DIGIT_BITS = 8>
struct RDXSRT_GPUDataManager {             // gpu_radix_sort.h:42-166: pre-allocated temporary memory
  void* workspace = nullptr; size_t bytes = 0;
  explicit RDXSRT_GPUDataManager(unsigned long long key_count, int key_type = -1, int value_bytes = -1) {
    if (key_type < 0) key_type = b200_shim_key_type<KeyT>();
    if (value_bytes < 0) value_bytes = b200_shim_value_bytes<ValueT>();
    if (b200_msb_sort(nullptr, nullptr, key_count, nullptr, nullptr, key_type, value_bytes, nullptr, &bytes, nullptr, nullptr, nullptr) == 0)
      cudaMalloc(&workspace, bytes);
  }
  ~RDXSRT_GPUDataManager() { if (workspace) cudaFree(workspace); }
  template <typename T> static int b200_shim_key_type();
  template <typename T> static int b200_shim_value_bytes();
};

namespace b200shim {
template <typename T> struct key_type_of;
template <> struct key_type_of<unsigned int> { static constexpr int value = B200_KEY_U32; };
template <> struct key_type_of<int> { static constexpr int value = B200_KEY_I32; };
template <> struct key_type_of<float> { static constexpr int value = B200_KEY_F32; };
template <> struct key_type_of<unsigned long long> { static constexpr int value = B200_KEY_U64; };
template <> struct key_type_of<unsigned long> { static constexpr int value = sizeof(unsigned long) == 8 ? B200_KEY_U64 : B200_KEY_U32; };
template <> struct key_type_of<long long> { static constexpr int value = B200_KEY_I64; };
template <> struct key_type_of<long> { static constexpr int value = sizeof(long) == 8 ? B200_KEY_I64 : B200_KEY_I32; };
template <> struct key_type_of<double> { static constexpr int value = B200_KEY_F64; };
inline void die_on_error(int rc, const char* what) {
  if (rc != 0) { fprintf(stderr, "%s failed: %s\n", what, b200_error_string(rc)); exit(-1); }
}
template <typename V> struct value_bytes_of { static constexpr int value = (int)sizeof(V); };
template <> struct value_bytes_of<cub::NullType> { static constexpr int value = 0; };
}  // namespace b200shim

template <typename KeyT, typename ValueT, typename IndexT, unsigned TPB, unsigned KPT, unsigned DIGIT_BITS>
template <typename T> int RDXSRT_GPUDataManager<KeyT, ValueT, IndexT, TPB, KPT, DIGIT_BITS>::b200_shim_key_type() { return b200shim::key_type_of<T>::value; }
template <typename KeyT, typename ValueT, typename IndexT, unsigned TPB, unsigned KPT, unsigned DIGIT_BITS>
template <typename T> int RDXSRT_GPUDataManager<KeyT, ValueT, IndexT, TPB, KPT, DIGIT_BITS>::b200_shim_value_bytes() { return b200shim::value_bytes_of<T>::value; }

// rdxsrt_unstable_sort (gpu_radix_sort.h:187-507): device pointers; both buffer pairs are clobbered; returns the buffers that
// hold the result.  Like the reference it is synchronous on return (gpu_radix_sort.h:489-491); unlike it, it is re-entrant.
template <typename KeyT, typename ValueT = cub::NullType, typename IndexT = unsigned int, unsigned int TPB = 0, unsigned int KPT = 0,
          unsigned int DIGIT_BITS = 8, unsigned int TINY_BUCKET_MERGE_THRESHOLD = 3000, unsigned int NUM_LSB_IN_VALUE = 0>
RDXSRT_SortedSequence<KeyT, ValueT> rdxsrt_unstable_sort(KeyT* dev_keys, ValueT* dev_values, IndexT key_count, KeyT* dev_sorted_keys_out,
                                                         ValueT* dev_sorted_values_out, LocalSortConfigSet<KeyT, ValueT>* = nullptr,
                                                         RDXSRT_GPUDataManager<KeyT, ValueT, IndexT, TPB, KPT, DIGIT_BITS>* pre_allocated_dm = nullptr,
                                                         cudaStream_t cstrm_extsrt = nullptr) {
  static_assert(NUM_LSB_IN_VALUE == 0, "the NUM_LSB_IN_VALUE extension (gpu_radix_sort.h:195) is not provided");
  constexpr int vb = b200shim::value_bytes_of<ValueT>::value;
  void* ok = dev_keys; void* ov = (void*)dev_values;
  size_t bytes = pre_allocated_dm ? pre_allocated_dm->bytes : 0;
  const int rc = b200_msb_sort(dev_keys, vb ? (void*)dev_values : nullptr, (uint64_t)key_count, dev_sorted_keys_out, vb ? (void*)dev_sorted_values_out : nullptr,
                               b200shim::key_type_of<KeyT>::value, vb, pre_allocated_dm ? pre_allocated_dm->workspace : nullptr,
                               pre_allocated_dm ? &bytes : nullptr, (b200_stream_t)cstrm_extsrt, &ok, &ov);
  b200shim::die_on_error(rc, "rdxsrt_unstable_sort");          // the reference exit(-1)s on failure (gpu_radix_sort.h:397-400); never return unsorted data
  b200shim::die_on_error((int)cudaStreamSynchronize(cstrm_extsrt), "rdxsrt_unstable_sort (synchronize)");
  if (pre_allocated_dm) {        // device-side conditions (a work list of the data manager's workspace overflowed): exit like the reference
    int status = 0;
    b200shim::die_on_error(b200_sort_status(pre_allocated_dm->workspace, (b200_stream_t)cstrm_extsrt, &status), "rdxsrt_unstable_sort (status)");
    if (status != 0) { fprintf(stderr, "rdxsrt_unstable_sort failed: device status %d\n", status); exit(-1); }
  }
  RDXSRT_SortedSequence<KeyT, ValueT> r;
  r.sorted_keys = (KeyT*)ok; r.sorted_values = (ValueT*)ov;
  return r;
}

// Host-pointer wrappers (gpu_radix_sort.h:510-541, 543-587).
template <typename KeyT>
void rdxsrt_unstable_sort_keys(KeyT* keys, unsigned long long key_count, KeyT* sorted_keys_out) {
  b200shim::die_on_error(b200_msb_sort_host(keys, nullptr, key_count, sorted_keys_out, nullptr, b200shim::key_type_of<KeyT>::value, 0), "rdxsrt_unstable_sort_keys");
}
template <typename KeyT, typename ValueT>
void rdxsrt_unstable_sort_pairs(KeyT* keys, ValueT* values, unsigned long long key_count, KeyT* sorted_keys_out, ValueT* sorted_values_out) {
  b200shim::die_on_error(b200_msb_sort_host(keys, values, key_count, sorted_keys_out, sorted_values_out, b200shim::key_type_of<KeyT>::value, (int)sizeof(ValueT)),
                         "rdxsrt_unstable_sort_pairs");
}
