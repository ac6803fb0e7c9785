// hist.cuh -- digit histograms.
//   hist_all_kernel : ALL digit histograms of an LSB sort in one read of the keys (replaces one upsweep per pass,
//                     lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:72-109).
//   seg_hist_kernel : one 256-bin histogram per MSB segment (replaces rdxsrt_histogram,
//                     msb/src/sort/cuda_radix_sort.h:657-802), driven by the device-built tile list.
//   scan_bins_kernel: exclusive scan of each 256-bin histogram -> absolute digit starts.
// Counting uses shared-memory atomics (result unused => RED), ~3500 G/s full chip on B200
// (profiles/ubench_rank_r01.jsonl), and stays fast when all digits are equal (ptxas aggregates the RED).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int HIST_THREADS = 512;
constexpr int MAX_PASSES = 8;

struct HistAllArgs {
  const void* keys; uint64_t n;
  int num_passes, begin_bit, end_bit;   // pass p covers bits [begin_bit + 8p, min(end_bit, begin_bit + 8p + 8))
  int tw_in; Twiddle tw;
  unsigned long long* hist;             // [num_passes][256], zeroed
};

template <typename K>
__global__ void __launch_bounds__(HIST_THREADS) hist_all_kernel(const __grid_constant__ HistAllArgs a) {
  __shared__ uint32_t sh[MAX_PASSES * RADIX];
  const K* __restrict__ keys = reinterpret_cast<const K*>(a.keys);
  for (int i = threadIdx.x; i < a.num_passes * RADIX; i += HIST_THREADS) sh[i] = 0;
  __syncthreads();
  constexpr int VEC = 16 / sizeof(K);
  auto count = [&](K k) {
    if (a.tw_in) k = twiddle_in<K>(k, a.tw);
#pragma unroll
    for (int p = 0; p < MAX_PASSES; ++p) {
      if (p < a.num_passes) {
        const int lo = a.begin_bit + 8 * p;
        const int nb = a.end_bit - lo < 8 ? a.end_bit - lo : 8;
        atomicAdd(&sh[p * RADIX + digit_of<K>(k, lo, (1u << nb) - 1u)], 1u);
      }
    }
  };
  // contiguous chunk per CTA with 16-byte vector loads when the pointer allows it
  const bool vec_ok = (reinterpret_cast<uintptr_t>(keys) & 15u) == 0;
  const uint64_t nvec = vec_ok ? a.n / VEC : 0;
  const uint64_t per = (nvec + gridDim.x - 1) / gridDim.x;
  const uint64_t v0 = per * blockIdx.x, v1 = v0 + per < nvec ? v0 + per : nvec;
  for (uint64_t v = v0 + threadIdx.x; v < v1; v += HIST_THREADS) {
    const uint4 q = reinterpret_cast<const uint4*>(keys)[v];
    if (sizeof(K) == 4) { count((K)q.x); count((K)q.y); count((K)q.z); count((K)q.w); }
    else { count((K)(((uint64_t)q.y << 32) | q.x)); count((K)(((uint64_t)q.w << 32) | q.z)); }
  }
  // scalar remainder (tail, or everything when the pointer is not 16-byte aligned)
  for (uint64_t i = nvec * VEC + (uint64_t)blockIdx.x * HIST_THREADS + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * HIST_THREADS)
    count(keys[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < a.num_passes * RADIX; i += HIST_THREADS) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&a.hist[i], (unsigned long long)c);
  }
}

// In-place exclusive scan of `rows` histograms of 256 bins, adding `base`.  One CTA of 256 threads per row.
static __global__ void __launch_bounds__(RADIX) scan_bins_kernel(unsigned long long* hist, uint64_t base) {
  __shared__ unsigned long long sh[RADIX];
  unsigned long long* h = hist + (uint64_t)blockIdx.x * RADIX;
  const unsigned t = threadIdx.x;
  const unsigned long long c = h[t];
  sh[t] = c;
  __syncthreads();
  for (int o = 1; o < RADIX; o <<= 1) {
    const unsigned long long v = t >= (unsigned)o ? sh[t - o] : 0ull;
    __syncthreads();
    sh[t] += v;
    __syncthreads();
  }
  h[t] = base + sh[t] - c;
}

// ---------------------------------------------------------------------------------------------------------
// MSB: per-segment histograms over the level's tile list.
// ---------------------------------------------------------------------------------------------------------
struct SegHistArgs {
  const void* keys;
  const Seg* segs; const TileDesc* descs; const uint32_t* num_tiles_ptr;
  uint32_t* seg_hist;                   // [segment][256], zeroed
  int tile;                             // keys per tile (same tiling as the partition kernel)
  int shift; int tw_in; Twiddle tw;
};

template <typename K>
__global__ void __launch_bounds__(HIST_THREADS) seg_hist_kernel(const __grid_constant__ SegHistArgs a) {
  __shared__ uint32_t sh[RADIX];
  const K* __restrict__ keys = reinterpret_cast<const K*>(a.keys);
  const uint32_t num_tiles = *a.num_tiles_ptr;
  const uint32_t per = (num_tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t t0 = per * blockIdx.x, t1 = min(t0 + per, num_tiles);
  if (t0 >= t1) return;
  if (threadIdx.x < RADIX) sh[threadIdx.x] = 0;
  __syncthreads();
  uint32_t cur_seg = a.descs[t0].seg;
  for (uint32_t t = t0; t < t1; ++t) {
    const TileDesc td = a.descs[t];
    if (td.seg != cur_seg) {            // block-uniform: flush the finished segment
      __syncthreads();
      if (threadIdx.x < RADIX) {
        const uint32_t c = sh[threadIdx.x];
        if (c) atomicAdd(&a.seg_hist[(uint64_t)cur_seg * RADIX + threadIdx.x], c);
        sh[threadIdx.x] = 0;
      }
      __syncthreads();
      cur_seg = td.seg;
    }
    const uint32_t cnt = td.cnt;
    const K* p = keys + td.off;
    for (uint32_t i = threadIdx.x; i < cnt; i += HIST_THREADS) {
      K k = p[i];
      if (a.tw_in) k = twiddle_in<K>(k, a.tw);
      const uint32_t d = digit_of<K>(k, a.shift, 0xFFu);
      atomicAdd(&sh[d], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < RADIX) {
    const uint32_t c = sh[threadIdx.x];
    if (c) atomicAdd(&a.seg_hist[(uint64_t)cur_seg * RADIX + threadIdx.x], c);
  }
}

}  // namespace b200
