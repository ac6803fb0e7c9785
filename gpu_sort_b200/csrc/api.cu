// api.cu -- the C ABI (include/b200sort.h): argument checks, key-type -> bit-transform mapping, dispatch to the
// per-(key width, value width) instantiations, and the host-pointer convenience wrappers.
#include <cuda_runtime.h>
#include <mutex>
#include "../../include/b200sort.h"
#include "common.cuh"
#include "sort_api.h"
#include "msb_sched.cuh"
#include <cstddef>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace b200;

// ---- per-kernel timing (sort_api.h: ProfScope) -----------------------------------------------------------------
namespace b200 {
int g_prof_enabled = 0;
int g_key_range_probe = 1;
unsigned long long g_prof_launches = 0;
namespace {
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr; cudaEventCreate(&e); return e;
}
}  // namespace
void prof_begin(const char* name, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  ProfRec r{name, prof_event(), prof_event()};
  cudaEventRecord(r.e0, s);
  g_prof_recs.push_back(r);
}
void prof_end(cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().e1, s);
}
}  // namespace b200

namespace b200 { size_t status_word_offset() { return offsetof(MsbCounters, error); } }

namespace {

bool make_twiddle(int key_type, int descending, Twiddle* tw, int* key_bytes) {
  Twiddle t{0, 0, 0};
  switch (key_type) {
    case B200_KEY_U32: *key_bytes = 4; break;
    case B200_KEY_U64: *key_bytes = 8; break;
    case B200_KEY_I32: *key_bytes = 4; t.sign_mask = 0x80000000ull; break;
    case B200_KEY_I64: *key_bytes = 8; t.sign_mask = 0x8000000000000000ull; break;
    case B200_KEY_F32: *key_bytes = 4; t.sign_mask = 0x80000000ull; t.float_mask = 0xFFFFFFFFull; break;
    case B200_KEY_F64: *key_bytes = 8; t.sign_mask = 0x8000000000000000ull; t.float_mask = ~0ull; break;
    default: return false;
  }
  if (descending) t.flip_mask = (*key_bytes == 4) ? 0xFFFFFFFFull : ~0ull;
  *tw = t;
  return true;
}

#define DISPATCH_KV(KB, VB, CALL)                                    \
  do {                                                               \
    if ((KB) == 4 && (VB) == 0) { using K = uint32_t; constexpr int V = 0; return (int)(CALL); } \
    if ((KB) == 4 && (VB) == 4) { using K = uint32_t; constexpr int V = 4; return (int)(CALL); } \
    if ((KB) == 4 && (VB) == 8) { using K = uint32_t; constexpr int V = 8; return (int)(CALL); } \
    if ((KB) == 8 && (VB) == 0) { using K = uint64_t; constexpr int V = 0; return (int)(CALL); } \
    if ((KB) == 8 && (VB) == 4) { using K = uint64_t; constexpr int V = 4; return (int)(CALL); } \
    if ((KB) == 8 && (VB) == 8) { using K = uint64_t; constexpr int V = 8; return (int)(CALL); } \
    return (int)cudaErrorInvalidValue;                               \
  } while (0)

// grow-only device scratch for the host-pointer wrappers (one set per process; serialised by a mutex)
struct HostPathCache {
  std::mutex mu;
  void* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[5] = {0, 0, 0, 0, 0};
  cudaStream_t stream = nullptr;
  cudaError_t need(int i, size_t bytes) {
    if (bytes <= cap[i]) return cudaSuccess;
    if (buf[i]) cudaFree(buf[i]);
    buf[i] = nullptr; cap[i] = 0;
    cudaError_t e = cudaMalloc(&buf[i], bytes);
    if (e == cudaSuccess) cap[i] = bytes;
    return e;
  }
};
HostPathCache g_caches[64];          // one per device: a process may drive several GPUs

int sort_host(bool msb, const void* hk, const void* hv, uint64_t n, void* hko, void* hvo, int key_type, int value_bytes, int descending) {
  Twiddle tw; int kb;
  if (!make_twiddle(key_type, descending, &tw, &kb)) return (int)cudaErrorInvalidValue;
  if (value_bytes != 0 && value_bytes != 4 && value_bytes != 8) return (int)cudaErrorInvalidValue;
  if ((value_bytes != 0) != (hv != nullptr)) return (int)cudaErrorInvalidValue;
  if (n == 0) return 0;
  int dev = 0;
  cudaGetDevice(&dev);
  HostPathCache& g_cache = g_caches[(dev >= 0 && dev < 64) ? dev : 0];
  std::lock_guard<std::mutex> lock(g_cache.mu);
  cudaError_t e;
  if (!g_cache.stream && (e = cudaStreamCreateWithFlags(&g_cache.stream, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
  cudaStream_t s = g_cache.stream;
  size_t ws = 0;
  if (msb) e = (cudaError_t)b200_msb_sort(nullptr, nullptr, n, nullptr, nullptr, key_type, value_bytes, nullptr, &ws, s, nullptr, nullptr);
  else e = (cudaError_t)b200_lsb_sort(nullptr, &ws, nullptr, nullptr, nullptr, nullptr, nullptr, n, key_type, value_bytes, 0, kb * 8, descending, 1, s);
  if (e != cudaSuccess) return (int)e;
  if ((e = g_cache.need(0, n * kb)) != cudaSuccess) return (int)e;
  if ((e = g_cache.need(1, n * kb)) != cudaSuccess) return (int)e;
  if (value_bytes) {
    if ((e = g_cache.need(2, n * value_bytes)) != cudaSuccess) return (int)e;
    if ((e = g_cache.need(3, n * value_bytes)) != cudaSuccess) return (int)e;
  }
  if ((e = g_cache.need(4, ws)) != cudaSuccess) return (int)e;
  void *k0 = g_cache.buf[0], *k1 = g_cache.buf[1], *v0 = value_bytes ? g_cache.buf[2] : nullptr, *v1 = value_bytes ? g_cache.buf[3] : nullptr;
  if ((e = cudaMemcpyAsync(k0, hk, n * kb, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
  if (value_bytes && (e = cudaMemcpyAsync(v0, hv, n * value_bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
  void *rk = k0, *rv = v0;
  if (msb) {
    e = (cudaError_t)b200_msb_sort(k0, v0, n, k1, v1, key_type, value_bytes, g_cache.buf[4], &ws, s, &rk, &rv);
  } else {
    int sel = 0;
    e = (cudaError_t)b200_lsb_sort(g_cache.buf[4], &ws, k0, k1, v0, v1, &sel, n, key_type, value_bytes, 0, kb * 8, descending, 1, s);
    if (sel) { rk = k1; rv = v1; }
  }
  if (e != cudaSuccess) return (int)e;
  if ((e = cudaMemcpyAsync(hko, rk, n * kb, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
  if (value_bytes && (e = cudaMemcpyAsync(hvo, rv, n * value_bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(s);
}

}  // namespace

extern "C" {

int b200_version(void) { return 102; }

int b200_prof_enable(int enable) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1); }
  g_prof_recs.clear();
  g_prof_enabled = enable ? 1 : 0;
  g_prof_launches = 0;
  return 0;
}

unsigned long long b200_prof_launches(void) { return g_prof_launches; }

int b200_prof_report(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  std::vector<std::string> order;
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : g_prof_recs) {
    cudaError_t e = cudaEventSynchronize(r.e1);
    if (e != cudaSuccess) return (int)e;
    float ms = 0.f;
    if ((e = cudaEventElapsedTime(&ms, r.e0, r.e1)) != cudaSuccess) return (int)e;
    auto it = agg.find(r.name);
    if (it == agg.end()) { order.push_back(r.name); agg[r.name] = {1, (double)ms}; }
    else { it->second.first += 1; it->second.second += ms; }
  }
  std::string out;
  char line[160];
  for (auto& n : order) {
    snprintf(line, sizeof line, "%s %d %.6f\n", n.c_str(), agg[n].first, agg[n].second);
    out += line;
  }
  if (buf && cap) { size_t m = out.size() < cap - 1 ? out.size() : cap - 1; memcpy(buf, out.data(), m); buf[m] = 0; }
  for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1); }
  g_prof_recs.clear();
  return 0;
}

const char* b200_error_string(int err) { return cudaGetErrorString((cudaError_t)err); }

int b200_lsb_sort(void* d_temp, size_t* temp_bytes, void* d_keys_current, void* d_keys_alternate, void* d_values_current,
                  void* d_values_alternate, int* selector_out, uint64_t num_items, int key_type, int value_bytes,
                  int begin_bit, int end_bit, int descending, int allow_overwrite, b200_stream_t stream) {
  Twiddle tw; int kb;
  if (!make_twiddle(key_type, descending, &tw, &kb) || temp_bytes == nullptr) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_KV(kb, value_bytes, (lsb_sort_impl<K, V>(d_temp, temp_bytes, d_keys_current, d_keys_alternate, d_values_current,
                                                    d_values_alternate, selector_out, num_items, tw, begin_bit, end_bit,
                                                    allow_overwrite, s)));
}

int b200_segmented_sort(void* d_temp, size_t* temp_bytes, void* d_keys_current, void* d_keys_alternate, void* d_values_current,
                        void* d_values_alternate, int* selector_out, uint64_t num_items, uint32_t num_segments,
                        const void* d_begin_offsets, const void* d_end_offsets, int offset_bytes, int key_type, int value_bytes,
                        int begin_bit, int end_bit, int descending, int allow_overwrite, b200_stream_t stream) {
  Twiddle tw; int kb;
  if (!make_twiddle(key_type, descending, &tw, &kb) || temp_bytes == nullptr) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_KV(kb, value_bytes, (segmented_sort_impl<K, V>(d_temp, temp_bytes, d_keys_current, d_keys_alternate, d_values_current,
                                                          d_values_alternate, selector_out, num_items, num_segments, d_begin_offsets,
                                                          d_end_offsets, offset_bytes, tw, begin_bit, end_bit, allow_overwrite, s)));
}

int b200_msb_sort_bits(void* d_keys, void* d_values, uint64_t num_items, void* d_keys_alt, void* d_values_alt, int key_type,
                       int value_bytes, int begin_bit, int end_bit, void* d_workspace, size_t* workspace_bytes, b200_stream_t stream,
                       void** out_keys, void** out_values) {
  Twiddle tw; int kb;
  if (!make_twiddle(key_type, 0, &tw, &kb)) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (d_workspace == nullptr && workspace_bytes == nullptr) {
    // reference default: temporary memory lives and dies inside the call (stream-ordered here, no host sync)
    size_t ws = 0;
    int e = b200_msb_sort_bits(nullptr, nullptr, num_items, nullptr, nullptr, key_type, value_bytes, begin_bit, end_bit, nullptr, &ws, stream, nullptr, nullptr);
    if (e) return e;
    void* w = nullptr;
    cudaError_t ce = cudaMallocAsync(&w, ws, s);
    if (ce != cudaSuccess) return (int)ce;
    e = b200_msb_sort_bits(d_keys, d_values, num_items, d_keys_alt, d_values_alt, key_type, value_bytes, begin_bit, end_bit, w, &ws, stream, out_keys, out_values);
    ce = cudaFreeAsync(w, s);
    return e ? e : (int)ce;
  }
  if (workspace_bytes == nullptr) return (int)cudaErrorInvalidValue;
  DISPATCH_KV(kb, value_bytes, (msb_sort_impl<K, V>(d_keys, d_values, num_items, d_keys_alt, d_values_alt, tw, d_workspace,
                                                    workspace_bytes, s, out_keys, out_values, begin_bit, end_bit)));
}

int b200_msb_sort(void* d_keys, void* d_values, uint64_t num_items, void* d_keys_alt, void* d_values_alt, int key_type,
                  int value_bytes, void* d_workspace, size_t* workspace_bytes, b200_stream_t stream, void** out_keys,
                  void** out_values) {
  return b200_msb_sort_bits(d_keys, d_values, num_items, d_keys_alt, d_values_alt, key_type, value_bytes, 0, 64, d_workspace, workspace_bytes, stream,
                            out_keys, out_values);
}

int b200_host_cache_release(void) {
  int dev = 0;
  cudaGetDevice(&dev);
  HostPathCache& c = g_caches[(dev >= 0 && dev < 64) ? dev : 0];
  std::lock_guard<std::mutex> lock(c.mu);
  for (int i = 0; i < 5; ++i) { if (c.buf[i]) cudaFree(c.buf[i]); c.buf[i] = nullptr; c.cap[i] = 0; }
  return 0;
}

int b200_set_key_range_probe(int enable) {
  const int old = b200::g_key_range_probe;
  b200::g_key_range_probe = enable ? 1 : 0;
  return old;
}

int b200_sort_status(const void* d_temp, b200_stream_t stream, int* status) {
  if (d_temp == nullptr || status == nullptr) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  uint32_t err = 0;
  cudaError_t e = cudaMemcpyAsync(&err, reinterpret_cast<const char*>(d_temp) + b200::status_word_offset(), sizeof err, cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return (int)e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return (int)e;
  *status = (int)err;
  return 0;
}

int b200_msb_sort_host(const void* h_keys, const void* h_values, uint64_t num_items, void* h_sorted_keys, void* h_sorted_values,
                       int key_type, int value_bytes) {
  return sort_host(true, h_keys, h_values, num_items, h_sorted_keys, h_sorted_values, key_type, value_bytes, 0);
}
int b200_lsb_sort_host(const void* h_keys, const void* h_values, uint64_t num_items, void* h_sorted_keys, void* h_sorted_values,
                       int key_type, int value_bytes, int descending) {
  return sort_host(false, h_keys, h_values, num_items, h_sorted_keys, h_sorted_values, key_type, value_bytes, descending);
}

int b200_range_partition_to(void* d_temp, size_t* temp_bytes, const void* d_keys_in, const void* d_values_in, uint64_t num_items, int key_type,
                            int value_bytes, int bits, const uint32_t* d_splitters, int num_parts, const uint64_t* d_local_counts,
                            uint64_t* d_part_offsets, const uint64_t* d_dst_keys, const uint64_t* d_dst_values, const uint64_t* d_dst_base,
                            b200_stream_t stream) {
  Twiddle tw; int kb;
  if (!make_twiddle(key_type, 0, &tw, &kb) || temp_bytes == nullptr) return (int)cudaErrorInvalidValue;
  if (d_temp != nullptr && (d_dst_keys == nullptr || d_dst_base == nullptr || (value_bytes && d_dst_values == nullptr))) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_KV(kb, value_bytes, (range_partition_impl<K, V>(d_temp, temp_bytes, d_keys_in, d_values_in, nullptr, nullptr, num_items, tw, bits,
                                                           d_splitters, num_parts, d_local_counts, d_part_offsets, d_dst_keys, d_dst_values,
                                                           d_dst_base, s)));
}

int b200_exchange_hist(void* d_temp, size_t* temp_bytes, const void* d_keys_in, uint64_t num_items, int key_type, int value_bytes, int bucket_bits,
                       uint64_t* d_hist, b200_stream_t stream) {
  Twiddle tw; int kb;
  if (!make_twiddle(key_type, 0, &tw, &kb) || temp_bytes == nullptr) return (int)cudaErrorInvalidValue;
  if (d_temp != nullptr && d_hist == nullptr) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_KV(kb, value_bytes, (exchange_hist_impl<K, V>(d_temp, temp_bytes, d_keys_in, num_items, tw, bucket_bits, d_hist, s)));
}

int b200_exchange_scatter(void* d_temp, size_t* temp_bytes, const void* d_keys_in, const void* d_values_in, uint64_t num_items, int key_type,
                          int value_bytes, int bucket_bits, const uint64_t* d_count_matrix, int num_ranks, int rank, uint64_t capacity,
                          const uint64_t* d_dst_keys, const uint64_t* d_dst_values, uint64_t* d_seg_begin, uint64_t* d_seg_end, uint64_t* d_info,
                          b200_stream_t stream) {
  Twiddle tw; int kb;
  if (!make_twiddle(key_type, 0, &tw, &kb) || temp_bytes == nullptr) return (int)cudaErrorInvalidValue;
  if (d_temp != nullptr && (d_count_matrix == nullptr || d_dst_keys == nullptr || (value_bytes && d_dst_values == nullptr) || d_seg_begin == nullptr ||
                            d_seg_end == nullptr || d_info == nullptr)) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_KV(kb, value_bytes, (exchange_scatter_impl<K, V>(d_temp, temp_bytes, d_keys_in, d_values_in, num_items, tw, bucket_bits, d_count_matrix, num_ranks,
                                                            rank, capacity, d_dst_keys, d_dst_values, d_seg_begin, d_seg_end, d_info, s)));
}

int b200_range_partition(void* d_temp, size_t* temp_bytes, const void* d_keys_in, const void* d_values_in, void* d_keys_out,
                         void* d_values_out, uint64_t num_items, int key_type, int value_bytes, int bits,
                         const uint32_t* d_splitters, int num_parts, const uint64_t* d_local_counts, uint64_t* d_part_offsets,
                         b200_stream_t stream) {
  Twiddle tw; int kb;
  if (!make_twiddle(key_type, 0, &tw, &kb) || temp_bytes == nullptr) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DISPATCH_KV(kb, value_bytes, (range_partition_impl<K, V>(d_temp, temp_bytes, d_keys_in, d_values_in, d_keys_out, d_values_out,
                                                           num_items, tw, bits, d_splitters, num_parts, d_local_counts,
                                                           d_part_offsets, nullptr, nullptr, nullptr, s)));
}

}  // extern "C"
