// ref_lsb_wrap.cu -- extern "C" handles onto the UNMODIFIED reference LSB driver functions
// sortPairsGPU / sortKeysGPU (/root/reference/lsb/sort.cu:25-76), compiled from where they lie, plus the
// same cub::DeviceRadixSort call shape on other key types (the "u32-key variant").
// The vendored CUB 1.6.4 cannot assemble for sm_100 (non-.sync shfl), so -- as the reference itself would
// have to on this box -- it builds against the toolkit's CUB 2.8.2 (onesweep, Policy1000).
// Test/bench infrastructure only (oracle/_ref/).
#define main ref_lsb_driver_main
#include "sort.cu"
#undef main
#include <cstdint>

template <typename K, typename V>
static int cub_pairs(void* temp, size_t* temp_bytes, K* k0, K* k1, V* v0, V* v1, int n, int descending, int begin_bit, int end_bit, int* selector) {
  cub::DoubleBuffer<K> dk(k0, k1);
  cub::DoubleBuffer<V> dv(v0, v1);
  cudaError_t e = descending ? cub::DeviceRadixSort::SortPairsDescending(temp, *temp_bytes, dk, dv, n, begin_bit, end_bit)
                             : cub::DeviceRadixSort::SortPairs(temp, *temp_bytes, dk, dv, n, begin_bit, end_bit);
  if (selector) *selector = dk.selector;
  return (int)e;
}
template <typename K>
static int cub_keys(void* temp, size_t* temp_bytes, K* k0, K* k1, int n, int descending, int begin_bit, int end_bit, int* selector) {
  cub::DoubleBuffer<K> dk(k0, k1);
  cudaError_t e = descending ? cub::DeviceRadixSort::SortKeysDescending(temp, *temp_bytes, dk, n, begin_bit, end_bit)
                             : cub::DeviceRadixSort::SortKeys(temp, *temp_bytes, dk, n, begin_bit, end_bit);
  if (selector) *selector = dk.selector;
  return (int)e;
}

extern "C" {
// The reference driver's own functions, float keys + uint values (lsb/sort.cu:25,49).  Return ms.
float ref_lsb_sortPairsGPU(float* k, float* k_alt, unsigned int* v, unsigned int* v_alt, int n) {
  return sortPairsGPU(k, k_alt, v, v_alt, n, g_allocator);
}
float ref_lsb_sortKeysGPU(float* k, float* k_alt, int n) { return sortKeysGPU(k, k_alt, n, g_allocator); }

// key_type: 0=u32 1=u64 2=i32 3=i64 4=f32 5=f64 ; value_bytes 0|4|8.  DoubleBuffer call shape, temp two-phase.
int ref_lsb_cub_sort(void* temp, size_t* temp_bytes, void* k0, void* k1, void* v0, void* v1, int n,
                     int key_type, int value_bytes, int descending, int begin_bit, int end_bit, int* selector) {
#define KCASE(T)                                                                                                    \
  if (value_bytes == 0) return cub_keys<T>(temp, temp_bytes, (T*)k0, (T*)k1, n, descending, begin_bit, end_bit, selector); \
  if (value_bytes == 4) return cub_pairs<T, unsigned int>(temp, temp_bytes, (T*)k0, (T*)k1, (unsigned int*)v0, (unsigned int*)v1, n, descending, begin_bit, end_bit, selector); \
  if (value_bytes == 8) return cub_pairs<T, unsigned long long>(temp, temp_bytes, (T*)k0, (T*)k1, (unsigned long long*)v0, (unsigned long long*)v1, n, descending, begin_bit, end_bit, selector); \
  return -1;
  switch (key_type) {
    case 0: KCASE(unsigned int)
    case 1: KCASE(unsigned long long)
    case 2: KCASE(int)
    case 3: KCASE(long long)
    case 4: KCASE(float)
    case 5: KCASE(double)
  }
#undef KCASE
  return -1;
}
}
