// ref_msb_wrap.cu -- extern "C" handles onto the UNMODIFIED reference MSB hybrid radix sort
// (/root/reference/msb/src/sort/gpu_radix_sort.h:187-507 device entry, :510-587 host wrappers).
// Test/bench infrastructure: lets tests/ and bench.py (--impl reference) run the reference itself on
// the GPU box as (1) the parity oracle's pin and (2) the reference arm.  No reference source is copied;
// the headers are included from where they lie.
#include "sort/gpu_radix_sort.h"
#include <cstdint>

template <typename K, typename V>
static int run_dev(K* keys, V* vals, unsigned long long n, K* keys_alt, V* vals_alt, void** out_k, void** out_v) {
  RDXSRT_SortedSequence<K, V> r = rdxsrt_unstable_sort<K, V, unsigned int>(keys, vals, (unsigned int)n, keys_alt, vals_alt);
  if (out_k) *out_k = (void*)r.sorted_keys;
  if (out_v) *out_v = (void*)r.sorted_values;
  return (int)cudaGetLastError();
}

// The same call with the reference's own pre_allocated_dm parameter (gpu_radix_sort.h:196,224-228): one RDXSRT_GPUDataManager per
// (type, key count) is built on first use and reused, so the eight cudaMalloc/cudaFree pairs of the default path stay outside
// the timed call -- the protocol SURVEY.md section 8d asks for (the as-shipped call above is reported next to it).
template <typename K, typename V>
static int run_dev_dm(K* keys, V* vals, unsigned long long n, K* keys_alt, V* vals_alt, void** out_k, void** out_v) {
  typedef RDXSRT_GPUDataManager<K, V, unsigned int> DM;
  static DM* dm = NULL; static unsigned long long dm_n = 0;
  if (dm == NULL || dm_n != n) {
    delete dm;
    dm = new DM((unsigned int)n, LocalSortConfigSet<K, V>::GetDefaultConfigSet());
    dm_n = n;
  }
  RDXSRT_SortedSequence<K, V> r = rdxsrt_unstable_sort<K, V, unsigned int>(keys, vals, (unsigned int)n, keys_alt, vals_alt, NULL, dm);
  if (out_k) *out_k = (void*)r.sorted_keys;
  if (out_v) *out_v = (void*)r.sorted_values;
  return (int)cudaGetLastError();
}

extern "C" {
int ref_msb_sort_device_prealloc(void* keys, void* vals, unsigned long long n, void* keys_alt, void* vals_alt,
                                 int key_bits, int value_bytes, void** out_k, void** out_v) {
  if (key_bits == 32 && value_bytes == 0) return run_dev_dm<unsigned int, cub::NullType>((unsigned int*)keys, NULL, n, (unsigned int*)keys_alt, NULL, out_k, out_v);
  if (key_bits == 64 && value_bytes == 0) return run_dev_dm<unsigned long long, cub::NullType>((unsigned long long*)keys, NULL, n, (unsigned long long*)keys_alt, NULL, out_k, out_v);
  if (key_bits == 32 && value_bytes == 4) return run_dev_dm<unsigned int, unsigned int>((unsigned int*)keys, (unsigned int*)vals, n, (unsigned int*)keys_alt, (unsigned int*)vals_alt, out_k, out_v);
  if (key_bits == 64 && value_bytes == 8) return run_dev_dm<unsigned long long, unsigned long long>((unsigned long long*)keys, (unsigned long long*)vals, n, (unsigned long long*)keys_alt, (unsigned long long*)vals_alt, out_k, out_v);
  return -1;
}
// key_bits: 32|64, value_bytes: 0|4|8.  Device pointers.  Returns cudaError_t of the last launch.
int ref_msb_sort_device(void* keys, void* vals, unsigned long long n, void* keys_alt, void* vals_alt,
                        int key_bits, int value_bytes, void** out_k, void** out_v) {
  if (key_bits == 32 && value_bytes == 0) return run_dev<unsigned int, cub::NullType>((unsigned int*)keys, NULL, n, (unsigned int*)keys_alt, NULL, out_k, out_v);
  if (key_bits == 64 && value_bytes == 0) return run_dev<unsigned long long, cub::NullType>((unsigned long long*)keys, NULL, n, (unsigned long long*)keys_alt, NULL, out_k, out_v);
  if (key_bits == 32 && value_bytes == 4) return run_dev<unsigned int, unsigned int>((unsigned int*)keys, (unsigned int*)vals, n, (unsigned int*)keys_alt, (unsigned int*)vals_alt, out_k, out_v);
  if (key_bits == 32 && value_bytes == 8) return run_dev<unsigned int, unsigned long long>((unsigned int*)keys, (unsigned long long*)vals, n, (unsigned int*)keys_alt, (unsigned long long*)vals_alt, out_k, out_v);
  if (key_bits == 64 && value_bytes == 4) return run_dev<unsigned long long, unsigned int>((unsigned long long*)keys, (unsigned int*)vals, n, (unsigned long long*)keys_alt, (unsigned int*)vals_alt, out_k, out_v);
  if (key_bits == 64 && value_bytes == 8) return run_dev<unsigned long long, unsigned long long>((unsigned long long*)keys, (unsigned long long*)vals, n, (unsigned long long*)keys_alt, (unsigned long long*)vals_alt, out_k, out_v);
  return -1;
}
// Host-pointer wrappers (gpu_radix_sort.h:510-587): malloc + H2D + sort + D2H + free, as the reference tests call them.
int ref_msb_sort_keys_host(void* keys, unsigned long long n, void* sorted_out, int key_bits) {
  if (key_bits == 32) rdxsrt_unstable_sort_keys<unsigned int>((unsigned int*)keys, n, (unsigned int*)sorted_out);
  else if (key_bits == 64) rdxsrt_unstable_sort_keys<unsigned long long>((unsigned long long*)keys, n, (unsigned long long*)sorted_out);
  else return -1;
  return (int)cudaGetLastError();
}
int ref_msb_sort_pairs_host(void* keys, void* vals, unsigned long long n, void* sorted_k, void* sorted_v, int key_bits, int value_bytes) {
  if (key_bits == 32 && value_bytes == 4) rdxsrt_unstable_sort_pairs<unsigned int, unsigned int>((unsigned int*)keys, (unsigned int*)vals, n, (unsigned int*)sorted_k, (unsigned int*)sorted_v);
  else if (key_bits == 32 && value_bytes == 8) rdxsrt_unstable_sort_pairs<unsigned int, unsigned long long>((unsigned int*)keys, (unsigned long long*)vals, n, (unsigned int*)sorted_k, (unsigned long long*)sorted_v);
  else if (key_bits == 64 && value_bytes == 4) rdxsrt_unstable_sort_pairs<unsigned long long, unsigned int>((unsigned long long*)keys, (unsigned int*)vals, n, (unsigned long long*)sorted_k, (unsigned int*)sorted_v);
  else if (key_bits == 64 && value_bytes == 8) rdxsrt_unstable_sort_pairs<unsigned long long, unsigned long long>((unsigned long long*)keys, (unsigned long long*)vals, n, (unsigned long long*)sorted_k, (unsigned long long*)sorted_v);
  else return -1;
  return (int)cudaGetLastError();
}
}
