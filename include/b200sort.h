/*
 * b200sort.h -- C ABI of the B200-native radix-sort library (libb200sort.so).
 *
 * This is the drop-in boundary for the reference's two host entry-point families (anilshanbhag/gpu-sort); each
 * entry point cites the reference interface it replaces (paths relative to the reference tree).  The reference has
 * no FFI layer -- its boundary is C++ templates -- so include/b200sort_cub_shim.cuh (cub::DeviceRadixSort /
 * DeviceSegmentedRadixSort) and include/shim/sort/gpu_radix_sort.h (rdxsrt_unstable_sort) re-create those template names on
 * top of this ABI, and INTEGRATION.md shows how lsb/sort.cu, msb/src/test.cu and msb/tests are re-pointed at them.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns a cudaError_t value (0 = cudaSuccess) and never exits
 *     (the reference calls exit(-1), msb/src/sort/gpu_radix_sort.h:397-400);
 *   - all device work is enqueued on `stream`; no allocation happens inside a call when the caller supplies the
 *     temporary storage, and the host never waits in the middle of a sort (the reference MSB blocks the host several
 *     times per pass and cudaMallocs inside the call, gpu_radix_sort.h:224-228,387,489-491).  One exception, for
 *     num_items >= 2^22 outside CUDA-graph capture: after enqueuing the whole sort the call waits for a 24-byte
 *     read-back taken behind the first histogram (key-range probe), i.e. it returns while the
 *     device is still sorting but not before the first pass over the keys has finished.  A caller that needs strictly
 *     host-asynchronous calls (e.g. the stream waits on work this host thread has not submitted yet) turns the probe off with
 *     b200_set_key_range_probe(0); sorts then never wait, at the price of whole sweeps on leading digits all keys share;
 *   - item counts are 64-bit (reference: int / unsigned int, device_radix_sort.cuh:154, gpu_radix_sort.h:190);
 *   - key order is the reference's bit-transform order (cub::Traits<T>::TwiddleIn, lsb/cub/cub/util_type.cuh:966-1089):
 *     unsigned as is, signed with the sign bit flipped, floating point as -NaN < -inf < ... < -0.0 < +0.0 < ... < +inf < +NaN.
 */
#ifndef B200SORT_H_
#define B200SORT_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* b200_stream_t; /* == cudaStream_t */

typedef enum {
  B200_KEY_U32 = 0, B200_KEY_U64 = 1, B200_KEY_I32 = 2, B200_KEY_I64 = 3, B200_KEY_F32 = 4, B200_KEY_F64 = 5
} b200_key_type;

/* Library version (major*10000 + minor*100 + patch). */
B200_API int b200_version(void);

/* Human-readable text for a return code of this library (cudaGetErrorString). */
B200_API const char* b200_error_string(int err);

/* ------------------------------------------------------------------------------------------------------------------
 * Stable LSB radix sort.
 * Replaces cub::DeviceRadixSort::SortPairs / SortPairsDescending / SortKeys / SortKeysDescending,
 *   DoubleBuffer overloads      lsb/cub/cub/device/device_radix_sort.cuh:248-273, 424-449, 595-620, 754-781
 *   pointer (non-overwriting)   lsb/cub/cub/device/device_radix_sort.cuh:147-179, 328-360, 506-535, 670-698
 * as called by the reference driver, lsb/sort.cu:25-76 (sortPairsGPU / sortKeysGPU).
 *
 *   d_temp == NULL      : write the required temporary-storage size to *temp_bytes and do nothing else
 *                         (CUB's two-phase protocol, dispatch_radix_sort.cuh:846-850,1110-1111).
 *   d_keys_current      : input keys (DoubleBuffer::Current()); d_keys_alternate: the other buffer.
 *   d_values_*          : NULL (and value_bytes 0) for keys-only; value_bytes is 4 or 8 otherwise.
 *   selector_out        : 0 -> the result is in the *_current buffers, 1 -> in the *_alternate buffers
 *                         (DoubleBuffer::selector after the call, dispatch_radix_sort.cuh:1153-1159).
 *   [begin_bit,end_bit) : key bits that take part in the comparison (CUB default 0 .. 8*sizeof(key)).
 *   allow_overwrite     : 1 = DoubleBuffer semantics (both buffers may be written);
 *                         0 = pointer-overload semantics: input left untouched, result in *_alternate.
 * Equal keys keep their input order (stable), ascending or descending.
 * ------------------------------------------------------------------------------------------------------------------ */
B200_API int b200_lsb_sort(void* d_temp, size_t* temp_bytes,
                  void* d_keys_current, void* d_keys_alternate,
                  void* d_values_current, void* d_values_alternate,
                  int* selector_out, uint64_t num_items,
                  int key_type, int value_bytes, int begin_bit, int end_bit,
                  int descending, int allow_overwrite, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Segmented stable radix sort: every segment [d_begin_offsets[i], d_end_offsets[i]) of the key (and value) array is sorted on
 * its own, in ONE call.
 * Replaces cub::DeviceSegmentedRadixSort::SortPairs / SortPairsDescending / SortKeys / SortKeysDescending,
 *   pointer overloads       lsb/cub/cub/device/device_segmented_radix_sort.cuh:140-181, 338-379, 534-572, 716-754
 *   DoubleBuffer overloads  lsb/cub/cub/device/device_segmented_radix_sort.cuh:247-283, 445-481, 630-666, 808-844
 * (dispatch: DeviceSegmentedRadixSortKernel, lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:321-436).
 *
 *   num_segments / d_*_offsets : offsets live on the device; offset_bytes = 4 (CUB's `const int*`) or 8.  A segment with
 *                                end <= begin is empty.  The usual CSR form passes d_offsets and d_offsets + 1.
 *                                Segments must not overlap.  Elements outside every segment are unspecified in the
 *                                result buffer, as in CUB.
 *   everything else            : as b200_lsb_sort (two-phase temporary storage, selector_out, [begin_bit,end_bit), descending,
 *                                allow_overwrite).  num_items = length of the key array (< 2^32).
 * Stable inside every segment.
 * ------------------------------------------------------------------------------------------------------------------  * allow_overwrite bit 1 (value 2, keys-only): the caller vouches that keys which tie on [begin_bit, end_bit) are EQUAL keys (every
 * segment shares its remaining bits), so their order is free and the cheaper unstable engine may run.
 */
B200_API int b200_segmented_sort(void* d_temp, size_t* temp_bytes,
                        void* d_keys_current, void* d_keys_alternate,
                        void* d_values_current, void* d_values_alternate,
                        int* selector_out, uint64_t num_items, uint32_t num_segments,
                        const void* d_begin_offsets, const void* d_end_offsets, int offset_bytes,
                        int key_type, int value_bytes, int begin_bit, int end_bit,
                        int descending, int allow_overwrite, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Unstable MSB hybrid radix sort.
 * Replaces rdxsrt_unstable_sort<KeyT,ValueT,IndexT>(dev_keys, dev_values|NULL, key_count, dev_sorted_keys_out,
 *   dev_sorted_values_out|NULL, cfg, pre_allocated_dm, stream)         msb/src/sort/gpu_radix_sort.h:187-507
 * as called by msb/src/test.cu:53,55 and msb/tests/test_sort_keys.cu:106 / test_sort_pairs.cu:158.
 *
 *   d_keys / d_values         : input; d_keys_alt / d_values_alt: equally sized alternate buffers.  BOTH pairs are
 *                               clobbered, as in the reference.
 *   d_values == NULL          : keys-only (the reference passes cub::NullType + NULL).
 *   d_workspace == NULL and workspace_bytes != NULL : size query (mirrors pre_allocated_dm sizing,
 *                               RDXSRT_GPUDataManager, gpu_radix_sort.h:89-141).
 *   d_workspace == NULL and workspace_bytes == NULL : the library allocates and frees the workspace itself with
 *                               stream-ordered allocation (the reference's default behaviour, gpu_radix_sort.h:224-228).
 *   out_keys / out_values     : receive the pointers that hold the sorted result -- RDXSRT_SortedSequence
 *                               (gpu_radix_sort.h:505-506).  For 4- and 8-byte keys this is the INPUT buffer pair,
 *                               exactly like the reference (gpu_radix_sort.h:359-360).
 * The key sequence is fully sorted ascending; the (key,value) multiset is preserved; the order of values among equal
 * keys is unspecified (the reference reserves output chunks with atomicAdd, cuda_radix_sort.h:408-417).
 * ------------------------------------------------------------------------------------------------------------------ */
B200_API int b200_msb_sort(void* d_keys, void* d_values, uint64_t num_items,
                  void* d_keys_alt, void* d_values_alt,
                  int key_type, int value_bytes,
                  void* d_workspace, size_t* workspace_bytes, b200_stream_t stream,
                  void** out_keys, void** out_values);

/* b200_msb_sort restricted to key bits [begin_bit, end_bit) of the order-transformed key (the counterpart of the bit range of
 * cub::DeviceRadixSort, device_radix_sort.cuh:154-157).  The multi-GPU path uses it: after the key-range exchange the leading
 * bits of every key a GPU holds are known to be equal.  The result lands in the input buffers when ceil((end-begin)/8) is even. */
B200_API int b200_msb_sort_bits(void* d_keys, void* d_values, uint64_t num_items,
                       void* d_keys_alt, void* d_values_alt,
                       int key_type, int value_bytes, int begin_bit, int end_bit,
                       void* d_workspace, size_t* workspace_bytes, b200_stream_t stream,
                       void** out_keys, void** out_values);

/* ------------------------------------------------------------------------------------------------------------------
 * Host-pointer convenience wrappers (device buffers, H2D, sort, D2H; synchronous).  The device buffers are kept cached per device
 * between calls (grow-only: the reference pays cudaMalloc/cudaFree on every call); b200_host_cache_release() frees them.
 * Replace rdxsrt_unstable_sort_keys / rdxsrt_unstable_sort_pairs  msb/src/sort/gpu_radix_sort.h:510-541, 543-587
 * and give the LSB path the same shape.  h_* may be pageable or pinned host memory; outputs may alias inputs.
 * ------------------------------------------------------------------------------------------------------------------ */
B200_API int b200_msb_sort_host(const void* h_keys, const void* h_values, uint64_t num_items,
                       void* h_sorted_keys, void* h_sorted_values, int key_type, int value_bytes);
B200_API int b200_lsb_sort_host(const void* h_keys, const void* h_values, uint64_t num_items,
                       void* h_sorted_keys, void* h_sorted_values, int key_type, int value_bytes, int descending);

/* Process-wide switch of the key-range probe (see "Conventions" above): 0 = no sort call ever waits on the host; returns the
 * previous setting.  Default 1. */
B200_API int b200_set_key_range_probe(int enable);

/* Device-side status of the last sort that used d_temp (b200_lsb_sort / b200_segmented_sort temporary storage or a
 * b200_msb_sort workspace): 0 = ok; bit 0 = segment list overflow or a segment offset outside [0, num_items] (the segment was
 * dropped), bit 1 = an on-chip work list overflowed, bit 2 = tile list overflow.  A non-zero status means the output is not
 * completely sorted (the reference exits the process in the corresponding situations, msb/src/sort/gpu_radix_sort.h:397-400).
 * Copies one word back and synchronises `stream`. */
B200_API int b200_sort_status(const void* d_temp, b200_stream_t stream, int* status);

/* Frees the device buffers the host-pointer wrappers keep cached on the current device between calls. */
B200_API int b200_host_cache_release(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU building blocks (one process per GPU; the collectives themselves are issued by the caller, see
 * gpu_sort_b200/dist.py).  No reference counterpart: the reference is single-GPU.
 *
 *   b200_msd_histogram  : counts[b] = number of keys whose top `bits` (1 <= bits <= 14) bits of the order-transformed key
 *                         equal b; counts is uint64[1 << bits] on the device and is overwritten.
 *   b200_range_partition: stable G-way partition of (keys, values) by destination rank, where the destination of a
 *                         key is the index of the first splitter strictly greater than its transformed top-`bits`
 *                         bucket: dest = #{ j : d_splitters[j] <= bucket }, splitters ascending, num_parts-1 of them.
 *                         d_local_counts = this rank's own b200_msd_histogram output (gives the part sizes without
 *                         another read of the keys).  d_part_offsets (uint64[num_parts+1], device) receives the start
 *                         of every part in the output.  Temporary storage follows the two-phase protocol of b200_lsb_sort.
 * ------------------------------------------------------------------------------------------------------------------ */
B200_API int b200_msd_histogram(const void* d_keys, uint64_t num_items, int key_type, int bits,
                       uint64_t* d_counts, b200_stream_t stream);
B200_API int b200_range_partition(void* d_temp, size_t* temp_bytes,
                         const void* d_keys_in, const void* d_values_in, void* d_keys_out, void* d_values_out,
                         uint64_t num_items, int key_type, int value_bytes, int bits,
                         const uint32_t* d_splitters, int num_parts, const uint64_t* d_local_counts,
                         uint64_t* d_part_offsets, b200_stream_t stream);
/* The same partition FUSED with the exchange: part j is written straight into its own destination buffer -- normally
 * the receive buffer of GPU j, mapped into this process (peer memory over NVLink) -- so the all-to-all happens inside the
 * scatter kernel's coalesced write-out, tile by tile, instead of as a separate collective.
 *   d_dst_keys / d_dst_values : uint64[num_parts] on the device, base ADDRESS of destination j's key / value buffer
 *   d_dst_base                : uint64[num_parts] on the device, index inside destination j's buffer where this rank's part
 *                               begins (sum of the counts of lower source ranks, from the all-gathered count matrix)
 * The caller synchronises the ranks after the kernel (all peers' stores must have landed before the local sort). */
B200_API int b200_range_partition_to(void* d_temp, size_t* temp_bytes,
                            const void* d_keys_in, const void* d_values_in,
                            uint64_t num_items, int key_type, int value_bytes, int bits,
                            const uint32_t* d_splitters, int num_parts, const uint64_t* d_local_counts,
                            uint64_t* d_part_offsets, const uint64_t* d_dst_keys, const uint64_t* d_dst_values,
                            const uint64_t* d_dst_base, b200_stream_t stream);

/* The exchange as LEVEL 0 of the sort (the default multi-GPU path, gpu_sort_b200/dist.py ExchangeSorter): every rank counts its
 * keys per tile and per leading 8-bit digit of the transformed key -- b200_exchange_hist, ONE read -- and reports how many fall
 * into each of the 2^bucket_bits exchange BUCKETS (bucket = leading bucket_bits bits, 0 <= bucket_bits <= 8; d_hist = uint64[256],
 * entries past 2^bucket_bits are zero).  The caller all-gathers the G x 256 matrix, and b200_exchange_scatter derives the same
 * plan on every rank (buckets dealt to the ranks in contiguous balanced groups; inside a destination buffer buckets ascend, each
 * bucket's keys in source-rank order) and runs the stable scatter whose per-bucket destinations are the peers' receive buffers.
 * Afterwards rank r holds whole buckets of the global sort: d_seg_begin / d_seg_end (uint64[256], device) bound them inside its
 * receive buffer, ready for b200_segmented_sort with end_bit = key bits - bucket_bits.  d_info (uint64[8], device): [0] keys received by this rank, [1] status
 * (1: some rank would receive more than `capacity` keys -- nothing was written, use the key-range partition above), [2] the
 * largest receive count, [3]/[4] first / one-past-last digit this rank owns.  Both calls share d_temp (same size, contents kept
 * from the first call to the second).  The caller synchronises the ranks around the scatter like b200_range_partition_to. */
B200_API int b200_exchange_hist(void* d_temp, size_t* temp_bytes, const void* d_keys_in, uint64_t num_items, int key_type,
                       int value_bytes, int bucket_bits, uint64_t* d_hist, b200_stream_t stream);
B200_API int b200_exchange_scatter(void* d_temp, size_t* temp_bytes, const void* d_keys_in, const void* d_values_in,
                          uint64_t num_items, int key_type, int value_bytes, int bucket_bits,
                          const uint64_t* d_count_matrix, int num_ranks, int rank, uint64_t capacity,
                          const uint64_t* d_dst_keys, const uint64_t* d_dst_values,
                          uint64_t* d_seg_begin, uint64_t* d_seg_end, uint64_t* d_info, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Benchmark / test utilities that run on the device (synthetic inputs of SURVEY.md section 8d and size-independent
 * result checks).  Same generator as oracle/radix_oracle.c (oracle_gen_key), so CPU and GPU see identical inputs.
 *   dist: 0 uniform | 1 AND of `param` streams (entropy levels of msb/tests/data_gen.h:44-76; param 0 = all zero)
 *         | 2 zipf-like, key = rank | 3 zipf-like, key = mix64(rank) | 4 presorted ascending | 5 descending | 6 constant
 * ------------------------------------------------------------------------------------------------------------------ */
B200_API int b200_util_generate_keys(void* d_keys, uint64_t num_items, uint64_t start_index, uint64_t total_items,
                            int key_bits, uint64_t seed, int dist, uint64_t param, b200_stream_t stream);
B200_API int b200_util_iota(void* d_values, uint64_t num_items, uint64_t start, int value_bytes, b200_stream_t stream);
/* d_out (uint64[4], device): [0] sum and [1] xor of a per-pair hash (order independent multiset digest, identical
 * to oracle_digest), [2] number of adjacent key pairs out of order under the transformed order, [3] number of
 * adjacent EQUAL-key pairs whose values descend (0 for a stable sort of iota values). */
B200_API int b200_util_check(const void* d_keys, const void* d_values, uint64_t num_items, int key_type, int value_bytes,
                    int descending, uint64_t* d_out, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Per-kernel timing for the benchmark's roofline figure (replaces the reference's BM_START/STOP_CUDA_EVENT macros
 * around histogram / pfx_sum / scatter / local_sort, msb/src/sort/gpu_radix_sort.h:266-269,278,284 and
 * msb/external/benchmark/benchmark.h:640-733).  While enabled, every kernel launch of the library is bracketed by CUDA
 * events on the launching stream.  b200_prof_report synchronises on them and writes one text line per kernel family,
 * "<name> <launches> <total ms>", then clears the records.  Off by default.
 * ------------------------------------------------------------------------------------------------------------------ */
B200_API int b200_prof_enable(int enable);
B200_API int b200_prof_report(char* buf, size_t buf_bytes);
/* Kernels the library launched since b200_prof_enable(1) (every launch site counts itself while profiling is on). */
B200_API unsigned long long b200_prof_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SORT_H_ */
