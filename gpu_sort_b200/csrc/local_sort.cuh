// local_sort.cuh -- small-bucket local sort: a bucket that fits one CTA's shared memory is finished on chip in a
// single read + write sweep.  Replaces do_locrec_radix_sort_keys (msb/src/sort/cuda_radix_sort.h:1332-1620) and,
// with `stable` set, DeviceRadixSortSingleTileKernel (lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:209-305).
//
// Persistent CTAs take the work items round-robin (items are independent); while one bucket is being sorted the
// next one is already being staged into the other shared-memory slot by a TMA bulk copy.  The result goes to the
// final buffer (in place when the bucket already lives there -- safe because the whole bucket is on chip before
// anything is stored).  Two algorithms:
//
//  * one-shot counting sort (unstable items, the MSB path): the top <= 16 of the remaining bits index 2^16 4-bit
//    shared-memory counters (32 KB -- this is what Blackwell's 228 KB of shared memory buys).  A key's atomicAdd
//    returns its rank inside its cell, a prefix sum over the counter words gives every cell's start, and the key goes
//    straight to start + rank: ONE ranking step instead of one per 8-bit digit.  If more than 16 bits remain, the
//    few keys that share a cell (cells hold <= 15 keys) are ordered by direct comparison inside the cell.  A cell
//    that would exceed 15 keys (heavy duplicates / low-entropy bits) sends the bucket to the generic path.
//  * generic LSD passes of 8 bits (any item; the only path for stable sorts): as in the reference
//    (cuda_radix_sort.h:1400-1481) the first pass may use the cheap unordered atomic ranking, later passes the
//    stable ranking of tile.cuh.
#pragma once
#include "async.cuh"
#include "tile.cuh"

namespace b200 {

struct LocalArgs {
  void* keys[3]; void* vals[3];          // the ping-pong buffers (LocalItem::src indexes them)
  void* keys_final; void* vals_final;
  const LocalItem* items; const uint32_t* num_items_ptr;
  int tw_in;                             // keys still in caller form (single-tile sorts)
  int tw_out;
  int stable;                            // every pass ordered
  int begin_bit;                         // lowest bit to sort (0 for MSB items)
  Twiddle tw;
};

constexpr int COUNT_MAX_BITS = 16;
template <int THREADS>
struct CountSmem {
  static constexpr int WORDS = (1 << COUNT_MAX_BITS) / 8;
  static constexpr int NSEG = (WORDS / 4 + THREADS - 1) / THREADS;   // scan segments of THREADS 4-word groups
  alignas(16) uint32_t nib[WORDS];     // 8 counters of 4 bits per word
  alignas(16) uint16_t wpre[WORDS];    // number of keys in all earlier words
  uint32_t wt[NSEG][32];               // per-warp totals of the scan segments
};

template <typename K, int VB, int THREADS, int IPT>
struct LocalSmem {
  static constexpr int CAP = THREADS * IPT;
  static constexpr int SLACK = 16 / sizeof(K);
  using V = typename ValType<VB>::type;
  static constexpr int VSLACK = 16 / sizeof(V);
  alignas(16) K stage[2][CAP + SLACK];
  alignas(16) V vstage[VB ? 2 : 1][VB ? CAP + VSLACK : 1];
  union {
    RankSmem<THREADS, true> ordered;
    RankSmem<THREADS, false> unordered;
    CountSmem<THREADS> count;
  } rank;
  alignas(8) uint64_t bar[2];
  LocalItem item[2];
  uint32_t skew[2], vskew[2];
};

// ---------------------------------------------------------------------------------------------------------------
// One-shot counting sort of the bucket staged at sk[skew .. skew+cnt) (values at sv[vskew ..]).
// ROWS = compile-time bound on keys per thread (the caller picks the smallest instantiation that covers cnt).
// Returns false (nothing modified) if a cell overflowed; on success the sorted bucket is at sk[0..cnt), sv[0..cnt).
// `after_count` runs once between the counting and the scatter (the producer thread issues its prefetch there).
// ---------------------------------------------------------------------------------------------------------------
template <typename K, int VB, int THREADS, int ROWS, typename AfterCount>
__device__ __forceinline__ bool count_sort_item(K* __restrict__ sk, typename ValType<VB>::type* __restrict__ sv, uint32_t skew,
                                                uint32_t vskew, uint32_t cnt, int lo, int hi, bool tw_in, const Twiddle& tw,
                                                CountSmem<THREADS>& cs, AfterCount after_count) {
  using V = typename ValType<VB>::type;
  constexpr int NSEG = CountSmem<THREADS>::NSEG;
  constexpr int NWARPS = THREADS / 32;
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const int cbits = hi - lo < COUNT_MAX_BITS ? hi - lo : COUNT_MAX_BITS;    // bits that index the counters
  const int vshift = hi - cbits;
  const uint32_t vmask = (1u << cbits) - 1u;
  const uint32_t words = cbits > 5 ? 1u << (cbits - 3) : 4u;
  const uint32_t groups = words / 4;

  for (uint32_t i = tid; i < groups; i += THREADS) reinterpret_cast<uint4*>(cs.nib)[i] = make_uint4(0, 0, 0, 0);
  K key[ROWS]; uint32_t pos[ROWS];
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    const uint32_t idx = j * THREADS + tid;
    K k = (K)0;
    if (idx < cnt) { k = sk[skew + idx]; if (tw_in) k = twiddle_in<K>(k, tw); }
    key[j] = k;
  }
  __syncthreads();
  int ovf = 0;
#pragma unroll
  for (int j = 0; j < ROWS; ++j)
    if ((uint32_t)(j * THREADS) + tid < cnt) {
      const uint32_t v = (uint32_t)(key[j] >> vshift) & vmask;
      const uint32_t sh = (v & 7u) * 4u;
      const uint32_t old = atomicAdd(&cs.nib[v >> 3], 1u << sh);
      pos[j] = (old >> sh) & 15u;
      ovf |= (pos[j] == 15u);
    }
  if (__syncthreads_or(ovf)) return false;
  after_count();

  // ---- exclusive prefix over the counter words; thread t owns the NSEG consecutive 4-word groups t*NSEG ..
  uint32_t tsum = 0;
#pragma unroll
  for (int k = 0; k < NSEG; ++k) {
    const uint32_t gi = tid * NSEG + k;
    if (gi < groups) {
      const uint4 q = reinterpret_cast<const uint4*>(cs.nib)[gi];
      // nibble sums of the four words at once: byte lanes hold <= 4 * 30
      uint32_t t = (q.x & 0x0F0F0F0Fu) + ((q.x >> 4) & 0x0F0F0F0Fu) + (q.y & 0x0F0F0F0Fu) + ((q.y >> 4) & 0x0F0F0F0Fu) +
                   (q.z & 0x0F0F0F0Fu) + ((q.z >> 4) & 0x0F0F0F0Fu) + (q.w & 0x0F0F0F0Fu) + ((q.w >> 4) & 0x0F0F0F0Fu);
      tsum += (t * 0x01010101u) >> 24;
    }
  }
  uint32_t inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (unsigned)o) inc += t;
  }
  if (lane == 31) cs.wt[0][w] = inc;
  __syncthreads();
  {
    const uint32_t wv = lane < (unsigned)NWARPS ? cs.wt[0][lane] : 0u;
    uint32_t wi = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= (unsigned)o) wi += t;
    }
    uint32_t run = __shfl_sync(0xffffffffu, wi - wv, w) + inc - tsum;
#pragma unroll
    for (int k = 0; k < NSEG; ++k) {
      const uint32_t gi = tid * NSEG + k;
      if (gi < groups) {
        const uint4 q = reinterpret_cast<const uint4*>(cs.nib)[gi];
        const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
        uint32_t p[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          p[e] = run;
          uint32_t t = (qq[e] & 0x0F0F0F0Fu) + ((qq[e] >> 4) & 0x0F0F0F0Fu);
          run += (t * 0x01010101u) >> 24;
        }
        reinterpret_cast<uint2*>(cs.wpre)[gi] = make_uint2(p[0] | (p[1] << 16), p[2] | (p[3] << 16));
      }
    }
  }
  __syncthreads();

  // ---- scatter to cell start + rank in cell
  V val[VB ? ROWS : 1];
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    const uint32_t idx = j * THREADS + tid;
    if (idx < cnt) {
      const uint32_t v = (uint32_t)(key[j] >> vshift) & vmask;
      const uint32_t sh = (v & 7u) * 4u;
      uint32_t below = cs.nib[v >> 3] & ((1u << sh) - 1u);
      below = (below & 0x0F0F0F0Fu) + ((below >> 4) & 0x0F0F0F0Fu);
      below = (below * 0x01010101u) >> 24;
      pos[j] += (uint32_t)cs.wpre[v >> 3] + below;
      if (VB) val[j] = sv[vskew + idx];
      sk[pos[j]] = key[j];
    }
  }
  __syncthreads();
  if (VB) {
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
      if ((uint32_t)(j * THREADS) + tid < cnt) sv[pos[j]] = val[j];
  }

  // ---- more than 16 bits left: order the keys that share a cell by direct comparison (cells hold <= 15 keys)
  if (hi - lo > COUNT_MAX_BITS) {
    __syncthreads();
    uint32_t npos[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      npos[j] = 0xFFFFFFFFu;
      if ((uint32_t)(j * THREADS) + tid < cnt) {
        const uint32_t v = (uint32_t)(key[j] >> vshift) & vmask;
        const uint32_t sh = (v & 7u) * 4u;
        const uint32_t wd = cs.nib[v >> 3];
        const uint32_t c = (wd >> sh) & 15u;
        if (c > 1) {
          uint32_t below = wd & ((1u << sh) - 1u);
          below = (below & 0x0F0F0F0Fu) + ((below >> 4) & 0x0F0F0F0Fu);
          below = (below * 0x01010101u) >> 24;
          const uint32_t cell0 = (uint32_t)cs.wpre[v >> 3] + below;
          const uint32_t r = pos[j] - cell0;
          uint32_t t = 0;
          for (uint32_t i = 0; i < c; ++i) {
            const K o = sk[cell0 + i];
            t += (o < key[j] || (o == key[j] && i < r)) ? 1u : 0u;
          }
          if (t != r) npos[j] = cell0 + t;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
      if (npos[j] != 0xFFFFFFFFu) {
        sk[npos[j]] = key[j];
        if (VB) sv[npos[j]] = val[j];
      }
  }
  __syncthreads();
  return true;
}

template <typename K, int VB, int THREADS, int IPT>
__global__ void __launch_bounds__(THREADS, (THREADS <= 512 ? 2 : 1)) local_sort_kernel(const __grid_constant__ LocalArgs a) {
  using V = typename ValType<VB>::type;
  using SM = LocalSmem<K, VB, THREADS, IPT>;
  constexpr unsigned PRODUCER = THREADS - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const uint32_t num_items = *a.num_items_ptr;
  K* __restrict__ keys_out = reinterpret_cast<K*>(a.keys_final);
  V* __restrict__ vals_out = reinterpret_cast<V*>(a.vals_final);

  auto stage_item = [&](int slot, uint32_t i, const LocalItem& it) {
    if (i < num_items) {
      const BulkWindow<K> bw(reinterpret_cast<const K*>(a.keys[it.src]), it.off, it.cnt);
      uint32_t bytes = bw.bytes;
      sm.skew[slot] = bw.skew;
      fence_proxy_async();
      if (VB) {
        const BulkWindow<V> vw(reinterpret_cast<const V*>(a.vals[it.src]), it.off, it.cnt);
        sm.vskew[slot] = vw.skew;
        bytes += vw.bytes;
        mbar_expect_tx(&sm.bar[slot], bytes);
        bulk_g2s(&sm.vstage[VB ? slot : 0][0], vw.src, vw.bytes, &sm.bar[slot]);
      } else {
        mbar_expect_tx(&sm.bar[slot], bytes);
      }
      bulk_g2s(&sm.stage[slot][0], bw.src, bw.bytes, &sm.bar[slot]);
      sm.item[slot] = it;
    } else {
      LocalItem none{}; none.cnt = 0xFFFFFFFFu;      // end marker
      sm.item[slot] = none;
    }
  };

  if (tid == PRODUCER) {
    mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
    LocalItem it0{};
    if (blockIdx.x < num_items) it0 = a.items[blockIdx.x];
    stage_item(0, blockIdx.x, it0);
  }
  __syncthreads();

  for (uint32_t iter = 0;; ++iter) {
    const int slot = (int)(iter & 1u);
    const LocalItem it = sm.item[slot];
    if (it.cnt == 0xFFFFFFFFu) break;
    const uint32_t cnt = it.cnt;
    const uint32_t skew = sm.skew[slot], vskew = VB ? sm.vskew[slot] : 0;
    const int rows = (int)((cnt + THREADS - 1) / THREADS);
    const int lo = a.begin_bit, hi = it.nbits;
    const int passes = hi > lo ? (hi - lo + 7) / 8 : 0;
    const bool first_ordered = a.stable != 0;

    // producer: fetch the next work item's descriptor (consumed when the prefetch is issued)
    const uint32_t next_i = blockIdx.x + (iter + 1) * gridDim.x;
    LocalItem it_next{};
    if (tid == PRODUCER && next_i < num_items) it_next = a.items[next_i];
    bool staged_next = false;
    auto prefetch = [&]() {
      if (tid == PRODUCER && !staged_next) stage_item(slot ^ 1, next_i, it_next);
      staged_next = true;
    };

    mbar_wait(&sm.bar[slot], (iter >> 1) & 1u);
    K* __restrict__ sk = &sm.stage[slot][0];
    V* __restrict__ sv = &sm.vstage[VB ? slot : 0][0];

    bool sorted = false;
    if (!first_ordered && hi - lo > 8) {
      // the other slot is free (its bucket was written out last iteration): the prefetch goes out after the count
      if (rows * 2 <= IPT) sorted = count_sort_item<K, VB, THREADS, IPT / 2>(sk, sv, skew, vskew, cnt, lo, hi, a.tw_in != 0, a.tw, sm.rank.count, prefetch);
      else if (rows * 4 <= IPT * 3) sorted = count_sort_item<K, VB, THREADS, IPT * 3 / 4>(sk, sv, skew, vskew, cnt, lo, hi, a.tw_in != 0, a.tw, sm.rank.count, prefetch);
      else sorted = count_sort_item<K, VB, THREADS, IPT>(sk, sv, skew, vskew, cnt, lo, hi, a.tw_in != 0, a.tw, sm.rank.count, prefetch);
    }

    if (!sorted) {
      // ---- generic path: shared memory -> registers (warp-contiguous layout when the first pass is ordered)
      K key[IPT]; V val[VB ? IPT : 1];
      uint32_t valid = 0;
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        K k = (K)~(K)0;
        if (j < rows) {
          const uint32_t idx = first_ordered ? (w * (uint32_t)rows * 32 + j * 32 + lane) : (j * THREADS + tid);
          if (idx < cnt) {
            k = sk[skew + idx];
            if (a.tw_in) k = twiddle_in<K>(k, a.tw);
            if (VB) val[j] = sv[vskew + idx];
            valid |= 1u << j;
          }
        }
        key[j] = k;
      }
      if (passes == 0) {      // nothing to sort: (twiddled) copy through the slot
        __syncthreads();
#pragma unroll
        for (int j = 0; j < IPT; ++j)
          if ((valid >> j) & 1u) {
            const uint32_t idx = first_ordered ? (w * (uint32_t)rows * 32 + j * 32 + lane) : (j * THREADS + tid);
            sk[idx] = key[j];
            if (VB) sv[idx] = val[j];
          }
        __syncthreads();
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
        // every thread has its keys in registers: the slot is reused as the exchange buffer (positions 0..cnt-1)
#pragma unroll
        for (int j = 0; j < IPT; ++j)
          if ((valid >> j) & 1u) {
            sk[pos[j]] = key[j];
            if (VB) sv[pos[j]] = val[j];
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
                k = sk[idx];
                if (VB) val[j] = sv[idx];
                valid |= 1u << j;
              }
            }
            key[j] = k;
          }
          // no barrier needed here: tile_positions() synchronises before anybody scatters again
        }
      }
    }
    prefetch();

    // ---- coalesced write-out of the sorted bucket
    for (uint32_t pidx = tid; pidx < cnt; pidx += THREADS) {
      K k = sk[pidx];
      if (a.tw_out) k = twiddle_out<K>(k, a.tw);
      keys_out[it.off + pidx] = k;
      if (VB) vals_out[it.off + pidx] = sv[pidx];
    }
    __syncthreads();
  }
}

}  // namespace b200
