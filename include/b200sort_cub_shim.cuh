// include/b200sort_cub_shim.cuh -- cub::DeviceRadixSort call shape on top of the C ABI (include/b200sort.h).
//
// Re-points the reference's LSB driver (lsb/sort.cu:25-76) at libb200sort.so without editing it:
//     nvcc ... -I <repo>/include/shim -I <repo>/include -I lsb -I lsb/cub/test lsb/sort.cu -L<repo>/gpu_sort_b200 -lb200sort -lcurand
// (include/shim/cub/device/device_radix_sort.cuh shadows the CUB header the driver includes and aliases cub::DeviceRadixSort
// to the struct below)
// Every `cub::DeviceRadixSort::SortPairs / SortKeys / ...Descending` in the translation unit then resolves to the struct
// below, which keeps CUB's signatures (lsb/cub/cub/device/device_radix_sort.cuh:147-179,248-273,328-360,424-449,506-535,
// 595-620,670-698,754-781): two-phase temporary storage, DoubleBuffer overloads (selector updated) and pointer overloads
// (input untouched), [begin_bit,end_bit), stream, debug_synchronous.  cub::DoubleBuffer itself comes from whichever CUB the
// translation unit already includes.
#pragma once
#include <cuda_runtime.h>
#include <cub/util_type.cuh>
#include "b200sort.h"

namespace cub {
namespace b200detail {
template <typename T> struct kt;
template <> struct kt<unsigned int> { static constexpr int v = B200_KEY_U32; };
template <> struct kt<int> { static constexpr int v = B200_KEY_I32; };
template <> struct kt<float> { static constexpr int v = B200_KEY_F32; };
template <> struct kt<unsigned long long> { static constexpr int v = B200_KEY_U64; };
template <> struct kt<unsigned long> { static constexpr int v = sizeof(unsigned long) == 8 ? B200_KEY_U64 : B200_KEY_U32; };
template <> struct kt<long long> { static constexpr int v = B200_KEY_I64; };
template <> struct kt<long> { static constexpr int v = sizeof(long) == 8 ? B200_KEY_I64 : B200_KEY_I32; };
template <> struct kt<double> { static constexpr int v = B200_KEY_F64; };
}  // namespace b200detail

struct B200DeviceRadixSort {
  template <typename KeyT, typename ValueT>
  static cudaError_t run(void* d_temp, size_t& bytes, KeyT* k_cur, KeyT* k_alt, ValueT* v_cur, ValueT* v_alt, int* selector, int n, int begin_bit,
                         int end_bit, bool descending, bool overwrite, cudaStream_t stream, bool debug_synchronous) {
    const int e = b200_lsb_sort(d_temp, &bytes, k_cur, k_alt, v_cur, v_alt, selector, (uint64_t)n, b200detail::kt<KeyT>::v,
                                v_cur || v_alt || !d_temp ? (int)sizeof(ValueT) * (std::is_same<ValueT, NullType>::value ? 0 : 1) : 0, begin_bit, end_bit,
                                descending ? 1 : 0, overwrite ? 1 : 0, (b200_stream_t)stream);
    if (e == 0 && debug_synchronous && d_temp) return cudaStreamSynchronize(stream);
    return (cudaError_t)e;
  }
#define B200_CUB_ENTRY(NAME, DESC)                                                                                                          \
  template <typename KeyT, typename ValueT>                                                                                                 \
  static cudaError_t NAME##Pairs##DESC(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>& d_values, \
                                       int num_items, int begin_bit = 0, int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0,           \
                                       bool debug_synchronous = false) {                                                                    \
    int sel = 0;                                                                                                                            \
    cudaError_t e = run<KeyT, ValueT>(d_temp_storage, temp_storage_bytes, d_keys.Current(), d_keys.Alternate(), d_values.Current(),         \
                                      d_values.Alternate(), &sel, num_items, begin_bit, end_bit, sizeof(#DESC) > 1, true, stream, debug_synchronous); \
    if (d_temp_storage && e == cudaSuccess) { d_keys.selector ^= sel; d_values.selector ^= sel; }                                           \
    return e;                                                                                                                               \
  }                                                                                                                                         \
  template <typename KeyT, typename ValueT>                                                                                                 \
  static cudaError_t NAME##Pairs##DESC(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,           \
                                       const ValueT* d_values_in, ValueT* d_values_out, int num_items, int begin_bit = 0,                   \
                                       int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0, bool debug_synchronous = false) {           \
    return run<KeyT, ValueT>(d_temp_storage, temp_storage_bytes, const_cast<KeyT*>(d_keys_in), d_keys_out, const_cast<ValueT*>(d_values_in), \
                             d_values_out, nullptr, num_items, begin_bit, end_bit, sizeof(#DESC) > 1, false, stream, debug_synchronous);     \
  }                                                                                                                                         \
  template <typename KeyT>                                                                                                                  \
  static cudaError_t NAME##Keys##DESC(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, int num_items,          \
                                      int begin_bit = 0, int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0,                           \
                                      bool debug_synchronous = false) {                                                                     \
    int sel = 0;                                                                                                                            \
    cudaError_t e = run<KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys.Current(), d_keys.Alternate(), nullptr, nullptr, &sel,   \
                                        num_items, begin_bit, end_bit, sizeof(#DESC) > 1, true, stream, debug_synchronous);                 \
    if (d_temp_storage && e == cudaSuccess) d_keys.selector ^= sel;                                                                         \
    return e;                                                                                                                               \
  }                                                                                                                                         \
  template <typename KeyT>                                                                                                                  \
  static cudaError_t NAME##Keys##DESC(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,            \
                                      int num_items, int begin_bit = 0, int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0,            \
                                      bool debug_synchronous = false) {                                                                     \
    return run<KeyT, NullType>(d_temp_storage, temp_storage_bytes, const_cast<KeyT*>(d_keys_in), d_keys_out, nullptr, nullptr, nullptr,     \
                               num_items, begin_bit, end_bit, sizeof(#DESC) > 1, false, stream, debug_synchronous);                         \
  }
  B200_CUB_ENTRY(Sort, )
  B200_CUB_ENTRY(Sort, Descending)
#undef B200_CUB_ENTRY
};

// cub::DeviceSegmentedRadixSort call shape (lsb/cub/cub/device/device_segmented_radix_sort.cuh:140-181,247-283,338-379,
// 445-481,534-572,630-666,716-754,808-844) on b200_segmented_sort: `const int*` segment offsets, otherwise as above.
struct B200DeviceSegmentedRadixSort {
  template <typename KeyT, typename ValueT>
  static cudaError_t run(void* d_temp, size_t& bytes, KeyT* k_cur, KeyT* k_alt, ValueT* v_cur, ValueT* v_alt, int* selector, int n, int num_segments,
                         const int* d_begin_offsets, const int* d_end_offsets, int begin_bit, int end_bit, bool descending, bool overwrite,
                         cudaStream_t stream, bool debug_synchronous) {
    const int e = b200_segmented_sort(d_temp, &bytes, k_cur, k_alt, v_cur, v_alt, selector, (uint64_t)n, (uint32_t)num_segments, d_begin_offsets,
                                      d_end_offsets, 4, b200detail::kt<KeyT>::v, std::is_same<ValueT, NullType>::value ? 0 : (int)sizeof(ValueT),
                                      begin_bit, end_bit, descending ? 1 : 0, overwrite ? 1 : 0, (b200_stream_t)stream);
    if (e == 0 && debug_synchronous && d_temp) return cudaStreamSynchronize(stream);
    return (cudaError_t)e;
  }
#define B200_CUB_SEG_ENTRY(DESC)                                                                                                            \
  template <typename KeyT, typename ValueT>                                                                                                 \
  static cudaError_t SortPairs##DESC(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, DoubleBuffer<ValueT>& d_values, \
                                     int num_items, int num_segments, const int* d_begin_offsets, const int* d_end_offsets, int begin_bit = 0, \
                                     int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0, bool debug_synchronous = false) {             \
    int sel = 0;                                                                                                                            \
    cudaError_t e = run<KeyT, ValueT>(d_temp_storage, temp_storage_bytes, d_keys.Current(), d_keys.Alternate(), d_values.Current(),         \
                                      d_values.Alternate(), &sel, num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit, end_bit, \
                                      sizeof(#DESC) > 1, true, stream, debug_synchronous);                                                  \
    if (d_temp_storage && e == cudaSuccess) { d_keys.selector ^= sel; d_values.selector ^= sel; }                                           \
    return e;                                                                                                                               \
  }                                                                                                                                         \
  template <typename KeyT, typename ValueT>                                                                                                 \
  static cudaError_t SortPairs##DESC(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,             \
                                     const ValueT* d_values_in, ValueT* d_values_out, int num_items, int num_segments,                      \
                                     const int* d_begin_offsets, const int* d_end_offsets, int begin_bit = 0, int end_bit = sizeof(KeyT) * 8, \
                                     cudaStream_t stream = 0, bool debug_synchronous = false) {                                             \
    return run<KeyT, ValueT>(d_temp_storage, temp_storage_bytes, const_cast<KeyT*>(d_keys_in), d_keys_out, const_cast<ValueT*>(d_values_in), \
                             d_values_out, nullptr, num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit, end_bit,            \
                             sizeof(#DESC) > 1, false, stream, debug_synchronous);                                                          \
  }                                                                                                                                         \
  template <typename KeyT>                                                                                                                  \
  static cudaError_t SortKeys##DESC(void* d_temp_storage, size_t& temp_storage_bytes, DoubleBuffer<KeyT>& d_keys, int num_items,            \
                                    int num_segments, const int* d_begin_offsets, const int* d_end_offsets, int begin_bit = 0,              \
                                    int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0, bool debug_synchronous = false) {              \
    int sel = 0;                                                                                                                            \
    cudaError_t e = run<KeyT, NullType>(d_temp_storage, temp_storage_bytes, d_keys.Current(), d_keys.Alternate(), nullptr, nullptr, &sel,   \
                                        num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit, end_bit, sizeof(#DESC) > 1, true, \
                                        stream, debug_synchronous);                                                                         \
    if (d_temp_storage && e == cudaSuccess) d_keys.selector ^= sel;                                                                         \
    return e;                                                                                                                               \
  }                                                                                                                                         \
  template <typename KeyT>                                                                                                                  \
  static cudaError_t SortKeys##DESC(void* d_temp_storage, size_t& temp_storage_bytes, const KeyT* d_keys_in, KeyT* d_keys_out,              \
                                    int num_items, int num_segments, const int* d_begin_offsets, const int* d_end_offsets,                  \
                                    int begin_bit = 0, int end_bit = sizeof(KeyT) * 8, cudaStream_t stream = 0,                             \
                                    bool debug_synchronous = false) {                                                                       \
    return run<KeyT, NullType>(d_temp_storage, temp_storage_bytes, const_cast<KeyT*>(d_keys_in), d_keys_out, nullptr, nullptr, nullptr,     \
                               num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit, end_bit, sizeof(#DESC) > 1, false,       \
                               stream, debug_synchronous);                                                                                  \
  }
  B200_CUB_SEG_ENTRY()
  B200_CUB_SEG_ENTRY(Descending)
#undef B200_CUB_SEG_ENTRY
};
}  // namespace cub
