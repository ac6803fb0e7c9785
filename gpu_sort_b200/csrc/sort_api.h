// sort_api.h -- declarations shared by the per-type instantiation units (inst.cu) and the C ABI (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

struct Twiddle;

// Optional per-kernel timing (b200_prof_enable / b200_prof_report in the C ABI): when enabled, every launch site is
// bracketed by two CUDA events on the launching stream.  Off by default -- a disabled scope costs one load and a branch.
extern int g_prof_enabled;
extern int g_key_range_probe;      // b200_set_key_range_probe: 0 = never wait on the host inside a sort call
size_t status_word_offset();      // where MsbCounters::error lives inside the temporary storage of the MSD engine
extern unsigned long long g_prof_launches;      // kernels launched while profiling is enabled (b200_prof_launches)
inline void note_launch() { if (g_prof_enabled) ++g_prof_launches; }
void prof_begin(const char* name, cudaStream_t s);
void prof_end(cudaStream_t s);
struct ProfScope {
  cudaStream_t s; bool on;
  ProfScope(const char* name, cudaStream_t stream) : s(stream), on(g_prof_enabled != 0) { if (on) prof_begin(name, s); }
  ~ProfScope() { if (on) prof_end(s); }
};

template <typename K, int VB>
cudaError_t lsb_sort_impl(void* d_temp, size_t* temp_bytes, void* k0, void* k1, void* v0, void* v1, int* selector,
                          uint64_t n, const Twiddle& tw, int begin_bit, int end_bit, int allow_overwrite, cudaStream_t s);

template <typename K, int VB>
cudaError_t segmented_sort_impl(void* d_temp, size_t* temp_bytes, void* k0, void* k1, void* v0, void* v1, int* selector,
                                uint64_t n, uint32_t num_segments, const void* d_begin, const void* d_end, int offset_bytes,
                                const Twiddle& tw, int begin_bit, int end_bit, int allow_overwrite, cudaStream_t s);

template <typename K, int VB>
cudaError_t msb_sort_impl(void* keys, void* vals, uint64_t n, void* keys_alt, void* vals_alt, const Twiddle& tw,
                          void* d_ws, size_t* ws_bytes, cudaStream_t s, void** out_keys, void** out_vals, int begin_bit, int end_bit);

template <typename K, int VB>
cudaError_t range_partition_impl(void* d_temp, size_t* temp_bytes, const void* kin, const void* vin, void* kout, void* vout,
                                 uint64_t n, const Twiddle& tw, int bits, const uint32_t* d_splitters, int num_parts,
                                 const uint64_t* d_local_counts, uint64_t* d_part_offsets, const uint64_t* d_dst_keys, const uint64_t* d_dst_vals,
                                 const uint64_t* d_dst_base, cudaStream_t s);

template <typename K, int VB>
cudaError_t exchange_hist_impl(void* d_temp, size_t* temp_bytes, const void* kin, uint64_t n, const Twiddle& tw, int bucket_bits, uint64_t* d_hist, cudaStream_t s);

template <typename K, int VB>
cudaError_t exchange_scatter_impl(void* d_temp, size_t* temp_bytes, const void* kin, const void* vin, uint64_t n, const Twiddle& tw, int bucket_bits,
                                  const uint64_t* d_matrix, int G, int rank, uint64_t cap, const uint64_t* d_dst_keys, const uint64_t* d_dst_vals,
                                  uint64_t* d_seg_begin, uint64_t* d_seg_end, uint64_t* d_info, cudaStream_t s);

}  // namespace b200
