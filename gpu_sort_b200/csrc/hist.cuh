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
// MSD levels: digit histogram of every TILE of the level (replaces rdxsrt_histogram + its per-block histograms,
// msb/src/sort/cuda_radix_sort.h:657-802, and the decoupled look-back of a onesweep pass: with the counts of every tile
// known before the scatter starts, a tile's destination is a pure function of three prefetchable loads).
//   tile_off[t][d]  = number of keys with digit d in the tiles of t's segment that precede t INSIDE t's group
//   group_tail[g][d]= the same running count at the end of group g (over the last segment the group touches)
//   group_flag[g]   = 1 if a segment starts inside group g
//   seg_hist[s][d]  = keys with digit d in segment s                                  (global atomics, zeroed)
// A group = HIST_GROUP consecutive entries of the tile list (segments are contiguous in the list); one CTA handles a
// whole group, so the running counts live in registers.  group_carry_kernel then chains the groups.
// ---------------------------------------------------------------------------------------------------------
#ifndef B200_HIST_GROUP
#define B200_HIST_GROUP 16
#endif
constexpr int HIST_GROUP = B200_HIST_GROUP;

struct TileHistArgs {
  const void* keys;
  const TileDesc* descs; const uint32_t* num_tiles_ptr;
  uint32_t* tile_off; uint32_t* group_tail; uint32_t* group_flag;
  uint16_t* tile_cnt;                          // optional: [tile][256] plain counts (for scatter_fast_kernel)
  uint32_t* seg_hist;
  int shift; uint32_t mask; int tw_in; Twiddle tw;
  const uint32_t* splitters; int num_parts;     // range mode: digit = #{ j < num_parts-1 : splitters[j] <= (key >> shift) }
  unsigned long long* key_or; unsigned long long* key_and;   // PROBE: OR / AND of all transformed keys are accumulated here
  uint32_t* ticket;                            // zeroed counter: groups are handed out dynamically (nullptr: round-robin)
  unsigned long long* seg_or; unsigned long long* seg_and;   // SEGP: per-SEGMENT OR / AND of the keys (initialised to 0 / ~0)
};

template <typename K, bool RANGE, bool PROBE, bool SEGP = false>
__global__ void __launch_bounds__(HIST_THREADS, sizeof(K) == 4 ? 4 : 2) tile_hist_kernel(const __grid_constant__ TileHistArgs a) {
  pdl_wait();
  __shared__ uint32_t sh[RADIX];
  __shared__ uint32_t s_sor[2], s_sand[2];     // SEGP: OR / AND of the current tile's keys (low, high word)
  K t_or = (K)0, t_and = (K)~(K)0;             // SEGP: this thread's share of them
  const K* __restrict__ keys = reinterpret_cast<const K*>(a.keys);
  const uint32_t num_tiles = *a.num_tiles_ptr;
  const uint32_t num_groups = (num_tiles + HIST_GROUP - 1) / HIST_GROUP;
  const unsigned tid = threadIdx.x;
  const int shift = a.shift; const uint32_t mask = a.mask;
  const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
  using S = typename std::make_signed<K>::type;
  __shared__ RangeLut rl;
  if (RANGE) { range_lut_build(rl, a.splitters, a.num_parts, (int)sizeof(K) * 8 - shift); __syncthreads(); }
  const int cshift = RANGE ? rl.cshift : 0;
  K acc_or = (K)0, acc_and = (K)~(K)0;
  // SUS (compile time): the tile belongs to a segment suspected of holding one repeated key (its parent's histogram had a single
  // non-empty bin): only then is the per-key OR / AND kept -- segments of ordinary inputs pay nothing for the check
  auto count = [&](K k, auto SUS) {
    if (a.tw_in) k = (K)(k ^ (((K)((S)k >> (sizeof(K) * 8 - 1)) & fl) | sg) ^ fp);
    if (PROBE) { acc_or |= k; acc_and &= k; }
    if (SEGP && decltype(SUS)::value) { t_or |= k; t_and &= k; }
    uint32_t d;
    if (!RANGE) d = digit_of<K>(k, shift, mask);
    else d = range_part(rl, (uint32_t)(k >> shift), cshift);
    atomicAdd(&sh[d], 1u);
  };
  __shared__ uint32_t s_next;
  for (uint32_t g = blockIdx.x;; g += gridDim.x) {
    if (a.ticket != nullptr) {       // dynamic hand-out: with ~3.5 groups per CTA a static deal leaves 13 % of the CTAs idle in the last round
      if (tid == 0) s_next = atomicAdd(a.ticket, 1u);
      __syncthreads();
      g = s_next;                    // (the barriers of the tile loop order this read before the next hand-out)
    }
    if (g >= num_groups) break;
    const uint32_t t0 = g * HIST_GROUP, t1 = min(t0 + HIST_GROUP, num_tiles);
    uint32_t run = 0, acc = 0, flag = 0, cur_seg = a.descs[t0].seg;      // digit owners (tid < 256)
    for (uint32_t t = t0; t < t1; ++t) {
      const TileDesc td = a.descs[t];
      if (tid < RADIX) sh[tid] = 0;
      if (SEGP) {
        if (tid == 0) { s_sor[0] = 0; s_sor[1] = 0; s_sand[0] = 0xFFFFFFFFu; s_sand[1] = 0xFFFFFFFFu; }
        t_or = (K)0; t_and = (K)~(K)0;
      }
      __syncthreads();
      const uint32_t cnt = td.cnt;
      const K* p = keys + td.off;
      // 16-byte vector loads over the aligned body of the tile, scalar head and tail
      constexpr uint32_t VEC = 16 / sizeof(K);
      const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(p) & 15u) / sizeof(K));
      const uint32_t head = min(cnt, mis ? VEC - mis : 0u);
      const uint32_t nvec = (cnt - head) / VEC;
      const uint4* pv = reinterpret_cast<const uint4*>(p + head);
      const bool sus = SEGP && td.pad != 0;
      auto count_tile = [&](auto SUS) {
        // batches of four independent 16-byte loads per thread before any counting (memory-level parallelism)
        for (uint32_t i0 = 0; i0 < nvec; i0 += 4 * HIST_THREADS) {
          uint4 q[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * HIST_THREADS + tid;
            q[u] = i < nvec ? pv[i] : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (i0 + u * HIST_THREADS + tid < nvec) {
              if (sizeof(K) == 4) { count((K)q[u].x, SUS); count((K)q[u].y, SUS); count((K)q[u].z, SUS); count((K)q[u].w, SUS); }
              else { count((K)(((uint64_t)q[u].y << 32) | q[u].x), SUS); count((K)(((uint64_t)q[u].w << 32) | q[u].z), SUS); }
            }
          }
        }
        if (tid < head) count(p[tid], SUS);
        const uint32_t tail0 = head + nvec * VEC;
        if (tail0 + tid < cnt) count(p[tail0 + tid], SUS);
      };
      if (sus) count_tile(std::true_type{}); else count_tile(std::false_type{});
      if (sus) {           // one REDUX pair per warp and key word, then one shared-memory atomic pair per warp
        const uint32_t o0 = __reduce_or_sync(0xffffffffu, (uint32_t)t_or), a0 = __reduce_and_sync(0xffffffffu, (uint32_t)t_and);
        uint32_t o1 = 0, a1 = 0xFFFFFFFFu;
        if (sizeof(K) == 8) {
          o1 = __reduce_or_sync(0xffffffffu, (uint32_t)((unsigned long long)t_or >> 32));
          a1 = __reduce_and_sync(0xffffffffu, (uint32_t)((unsigned long long)t_and >> 32));
        }
        if ((tid & 31u) == 0) { atomicOr(&s_sor[0], o0); atomicAnd(&s_sand[0], a0); if (sizeof(K) == 8) { atomicOr(&s_sor[1], o1); atomicAnd(&s_sand[1], a1); } }
      }
      __syncthreads();
      if (sus && tid == 0 && cnt != 0) {
        atomicOr(&a.seg_or[td.seg], ((unsigned long long)s_sor[1] << 32) | s_sor[0]);
        atomicAnd(&a.seg_and[td.seg], ((unsigned long long)s_sand[1] << 32) | s_sand[0]);
      }
      if (tid < RADIX) {
        const uint32_t c = sh[tid];
        if (td.tile_in_seg == 0) {                       // a segment starts here: close the previous one
          if (acc) atomicAdd(&a.seg_hist[(uint64_t)cur_seg * RADIX + tid], acc);
          acc = 0; run = 0; flag = 1; cur_seg = td.seg;
        }
        a.tile_off[(uint64_t)t * RADIX + tid] = run;
        if (a.tile_cnt != nullptr) a.tile_cnt[(uint64_t)t * RADIX + tid] = (uint16_t)c;
        run += c; acc += c;
      }
    }
    if (tid < RADIX) {
      if (acc) atomicAdd(&a.seg_hist[(uint64_t)cur_seg * RADIX + tid], acc);
      a.group_tail[(uint64_t)g * RADIX + tid] = run;
      if (tid == 0) a.group_flag[g] = flag;
    }
  }
  if (PROBE) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc_or |= (K)__shfl_xor_sync(0xffffffffu, acc_or, o);
      acc_and &= (K)__shfl_xor_sync(0xffffffffu, acc_and, o);
    }
    if ((tid & 31u) == 0) {
      atomicOr(a.key_or, (unsigned long long)acc_or);
      atomicAnd(a.key_and, (unsigned long long)acc_and | (sizeof(K) == 4 ? 0xFFFFFFFF00000000ull : 0ull));
    }
  }
}

// carry[g][d] = keys with digit d in the tiles of the segment that CONTINUES into group g from earlier groups, i.e. a
// segmented exclusive scan of group_tail along the group axis (state after group g: flag[g] ? tail[g] : state + tail[g]).
// Grid = 8 CTAs (one per slab of 32 digits, lane = digit, 128-byte rows); the 32 warps of a CTA split the group range.
constexpr int CARRY_WARPS = 32;
static __global__ void __launch_bounds__(CARRY_WARPS * 32) group_carry_kernel(const uint32_t* group_tail, const uint32_t* group_flag, uint32_t* carry,
                                                                              const uint32_t* num_tiles_ptr) {
  pdl_wait();
  __shared__ uint32_t p_sum[CARRY_WARPS][32];
  __shared__ uint32_t p_flag[CARRY_WARPS];
  const uint32_t num_groups = (*num_tiles_ptr + HIST_GROUP - 1) / HIST_GROUP;
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const unsigned d = blockIdx.x * 32 + lane;
  const uint32_t per = (num_groups + CARRY_WARPS - 1) / CARRY_WARPS;
  const uint32_t g0 = min(per * w, num_groups), g1 = min(g0 + per, num_groups);
  uint32_t sum = 0, flag = 0;
  // (eight independent loads in flight per lane: the chain of dependent L2 round trips was most of this kernel's 50 us)
  for (uint32_t gb = g0; gb < g1; gb += 8) {
    uint32_t t[8], f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t g = gb + u;
      t[u] = g < g1 ? group_tail[(uint64_t)g * RADIX + d] : 0u;
      f[u] = g < g1 ? group_flag[g] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (gb + u < g1) { sum = f[u] ? t[u] : sum + t[u]; flag |= f[u]; }
  }
  p_sum[w][lane] = sum;
  if (lane == 0) p_flag[w] = flag;
  __syncthreads();
  uint32_t state = 0;
  for (unsigned ww = 0; ww < w; ++ww) state = p_flag[ww] ? p_sum[ww][lane] : state + p_sum[ww][lane];
  for (uint32_t gb = g0; gb < g1; gb += 8) {
    uint32_t t[8], f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t g = gb + u;
      t[u] = g < g1 ? group_tail[(uint64_t)g * RADIX + d] : 0u;
      f[u] = g < g1 ? group_flag[g] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (gb + u < g1) {
        carry[(uint64_t)(gb + u) * RADIX + d] = state;
        state = f[u] ? t[u] : state + t[u];
      }
  }
}

}  // namespace b200
