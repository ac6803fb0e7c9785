// local_sort.cuh -- small-bucket local sort: a bucket that fits one CTA's registers + shared memory is finished
// on chip in a single read + write sweep.  Replaces do_locrec_radix_sort_keys
// (msb/src/sort/cuda_radix_sort.h:1332-1620) and, with `stable` set, DeviceRadixSortSingleTileKernel
// (lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:209-305).
//
// LSD passes of 8 bits over the item's remaining bits, keys (and values) ping-ponging registers <-> shared
// memory.  As in the reference (cuda_radix_sort.h:1400-1481) the first pass may use the cheap unordered
// shared-memory-atomic ranking; every later pass uses the stable ranking.  Persistent CTAs pull items from a
// ticket; the result goes to the final buffer (in place when the bucket already lives there -- safe because the
// whole bucket is loaded before anything is stored).
#pragma once
#include "tile.cuh"

namespace b200 {

struct LocalArgs {
  void* keys[2]; void* vals[2];          // the two ping-pong buffers
  void* keys_final; void* vals_final;
  const LocalItem* items; const uint32_t* num_items_ptr; uint32_t* ticket;
  int tw_in;                             // keys still in caller form (single-tile sorts)
  int tw_out;
  int stable;                            // every pass ordered
  int begin_bit;                         // lowest bit to sort (0 for MSB items)
  Twiddle tw;
};

template <typename K, int VB, int THREADS, int IPT>
struct LocalSmem {
  static constexpr int CAP = THREADS * IPT;
  using V = typename ValType<VB>::type;
  alignas(16) K keys[CAP];
  alignas(16) V vals[VB ? CAP : 1];
  union {
    RankSmem<THREADS, true> ordered;
    RankSmem<THREADS, false> unordered;
  } rank;
  uint32_t item;
};

template <typename K, int VB, int THREADS, int IPT>
__global__ void __launch_bounds__(THREADS, 2) local_sort_kernel(const __grid_constant__ LocalArgs a) {
  using V = typename ValType<VB>::type;
  using SM = LocalSmem<K, VB, THREADS, IPT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const uint32_t num_items = *a.num_items_ptr;
  K* __restrict__ keys_out = reinterpret_cast<K*>(a.keys_final);
  V* __restrict__ vals_out = reinterpret_cast<V*>(a.vals_final);

  while (true) {
    if (tid == 0) sm.item = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t item = sm.item;
    if (item >= num_items) break;
    const LocalItem it = a.items[item];
    const uint32_t cnt = it.cnt;
    const int rows = (int)((cnt + THREADS - 1) / THREADS);
    const K* __restrict__ kin = reinterpret_cast<const K*>(a.keys[it.src]) + it.off;
    const V* __restrict__ vin = reinterpret_cast<const V*>(a.vals[it.src]) + it.off;
    const int lo = a.begin_bit, hi = it.nbits;
    const int passes = hi > lo ? (hi - lo + 7) / 8 : 0;
    const bool first_ordered = a.stable != 0;

    // ---- load (warp-contiguous layout when the first pass is ordered, else block-striped)
    K key[IPT]; V val[VB ? IPT : 1];
    uint32_t valid = 0;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      K k = (K)~(K)0;
      if (j < rows) {
        const uint32_t idx = first_ordered ? (w * (uint32_t)rows * 32 + j * 32 + lane) : (j * THREADS + tid);
        if (idx < cnt) {
          k = kin[idx];
          if (a.tw_in) k = twiddle_in<K>(k, a.tw);
          if (VB) val[j] = vin[idx];
          valid |= 1u << j;
        }
      }
      key[j] = k;
    }

    if (passes == 0) {      // nothing to sort: (twiddled) copy
#pragma unroll
      for (int j = 0; j < IPT; ++j)
        if ((valid >> j) & 1u) {
          const uint32_t idx = first_ordered ? (w * (uint32_t)rows * 32 + j * 32 + lane) : (j * THREADS + tid);
          K k = key[j];
          if (a.tw_out) k = twiddle_out<K>(k, a.tw);
          keys_out[it.off + idx] = k;
          if (VB) vals_out[it.off + idx] = val[j];
        }
      __syncthreads();
      continue;
    }

    for (int p = 0; p < passes; ++p) {
      const int shift = lo + 8 * p;
      const int nb = hi - shift < 8 ? hi - shift : 8;
      const uint32_t mask = (1u << nb) - 1u;
      uint32_t pos[IPT], t0, t1;
      auto dfn = [&](K k) { return digit_of<K>(k, shift, mask); };
      if (p == 0 && !first_ordered)
        tile_positions<THREADS, IPT, false>(key, dfn, valid, rows, 0u, mask, pos, sm.rank.unordered, t0, t1);
      else
        tile_positions<THREADS, IPT, true>(key, dfn, valid, rows, (uint32_t)rows * THREADS - cnt, mask, pos, sm.rank.ordered, t0, t1);
#pragma unroll
      for (int j = 0; j < IPT; ++j)
        if ((valid >> j) & 1u) {
          sm.keys[pos[j]] = key[j];
          if (VB) sm.vals[pos[j]] = val[j];
        }
      __syncthreads();
      if (p + 1 < passes) {   // read back in warp-contiguous order for the next (ordered) pass
        valid = 0;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
          K k = (K)~(K)0;
          if (j < rows) {
            const uint32_t idx = w * (uint32_t)rows * 32 + j * 32 + lane;
            if (idx < cnt) {
              k = sm.keys[idx];
              if (VB) val[j] = sm.vals[idx];
              valid |= 1u << j;
            }
          }
          key[j] = k;
        }
        // no barrier needed here: tile_positions() synchronises before anybody scatters again
      }
    }

    // ---- coalesced write-out of the sorted bucket
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const uint32_t pidx = j * THREADS + tid;
      if (pidx < cnt) {
        K k = sm.keys[pidx];
        if (a.tw_out) k = twiddle_out<K>(k, a.tw);
        keys_out[it.off + pidx] = k;
        if (VB) vals_out[it.off + pidx] = sm.vals[pidx];
      }
    }
    __syncthreads();
  }
}

}  // namespace b200
