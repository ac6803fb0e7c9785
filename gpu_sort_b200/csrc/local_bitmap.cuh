// local_bitmap.cuh -- on-chip finish of keys-only buckets with at most 16 undecided bits: a presence-bitmap sort.
// Replaces do_locrec_radix_sort_keys (msb/src/sort/cuda_radix_sort.h:1332-1620) for the buckets that dominate a
// 4-byte keys-only sort (2^28 uniform keys: 65 536 buckets of ~4096 keys with 16 bits left).
//
// With <= 16 bits left a bucket of ~4096 keys touches ~6 % of its 65 536 possible values, so the sorted bucket is
// described completely by WHICH values occur (and how often):
//   phase A  every key sets its bit in a 65 536-bit presence map (one shared-memory atomicOr per key; the few keys that
//            find their bit already set (3 % for uniform keys) set it in a second map, third and later copies go to a
//            short list); the keys themselves are not needed again;
//   phase B  each thread owns 256 consecutive cells (8 words of either map): popc -> block scan -> where its values start;
//   phase C  each thread walks the set bits of its words and EMITS the keys (bucket prefix | cell) in order into the
//            consumed staging slot, laid out so that shared-memory and output 16-byte vectors coincide;
//   phase D  coalesced 16-byte stores.
// ~35 thread-instructions per key against ~112 for the 4-bit-counter counting sort it replaces (profiles/r01_full_cfg2.txt),
// no per-key rank lookup, no reorder pass.  Only valid where equal cells mean equal KEYS, i.e. the sort covers the
// whole key (begin_bit == 0, and every bit above the bucket's undecided bits is shared by the bucket): the host routes
// other windows to the counting / LSD kernels.  A bucket with more than BITMAP_EXTRA third-or-later copies (heavy
// duplicates) is handed to the overflow list that an ALGO_LSD launch finishes.
#pragma once
#include <cstring>
#include "async.cuh"
#include "common.cuh"
#include "local_sort.cuh"

namespace b200 {

constexpr int BITMAP_BITS = 16;
constexpr int BITMAP_WORDS = (1 << BITMAP_BITS) / 32;
constexpr int BITMAP_EXTRA = 64;

template <typename K, int THREADS, int CAP>
struct BitmapSmem {
  static constexpr int E = 16 / sizeof(K);
  static constexpr int WPT = BITMAP_WORDS / THREADS;        // words of either map a thread owns
  static_assert(BITMAP_WORDS % THREADS == 0 && WPT % 4 == 0, "a thread owns whole 16-byte vectors of the maps");
  alignas(16) K stage[2][CAP + 2 * E];
  alignas(16) uint32_t bm1[BITMAP_WORDS];       // value occurs at least once
  alignas(16) uint32_t bm2[BITMAP_WORDS];       // ... at least twice
  uint32_t extras[BITMAP_EXTRA];                // third and later copies (cell numbers)
  uint32_t nextra;
  uint32_t wt[32];
  alignas(8) uint64_t bar[2];
  LocalItem item[2];
  uint32_t skew[2];
};

template <typename K, int THREADS, int CAP, int OCC>
__global__ void __launch_bounds__(THREADS, OCC) bitmap_sort_kernel(const __grid_constant__ LocalArgs a) {
  pdl_wait();
  using SM = BitmapSmem<K, THREADS, CAP>;
  constexpr unsigned PRODUCER = THREADS - 1;
  constexpr int E = SM::E, WPT = SM::WPT, NWARPS = THREADS / 32;
  constexpr int MAXIT = ((CAP + 2 * E) / E + THREADS - 1) / THREADS;       // 16-byte vectors of a staged bucket per thread
  static_assert(MAXIT * E <= 32, "one duplicate flag per key of a thread");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const uint32_t num_items = min(*a.num_items_ptr, a.max_items);
  K* __restrict__ keys_out = reinterpret_cast<K*>(a.keys_final);

  auto stage_item = [&](int slot, uint32_t i, const LocalItem& it) {
    if (i < num_items) {
      const BulkWindow<K> bw(reinterpret_cast<const K*>(a.keys[it.src]), it.off, it.cnt);
      sm.skew[slot] = bw.skew;
      fence_proxy_async();
      mbar_expect_tx(&sm.bar[slot], bw.bytes);
      bulk_g2s(&sm.stage[slot][0], bw.src, bw.bytes, &sm.bar[slot]);
      sm.item[slot] = it;
    } else {
      LocalItem none{}; none.cnt = 0xFFFFFFFFu;
      sm.item[slot] = none;
    }
  };

  for (int i = tid; i < BITMAP_WORDS / 4; i += THREADS) {
    reinterpret_cast<uint4*>(sm.bm1)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(sm.bm2)[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) sm.nextra = 0;
  LocalItem it_a{}, it_b{};
  if (tid == PRODUCER) {
    mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
    LocalItem it0{};
    if (blockIdx.x < num_items) it0 = a.items[blockIdx.x];
    if (blockIdx.x + gridDim.x < num_items) it_a = a.items[blockIdx.x + gridDim.x];
    stage_item(0, blockIdx.x, it0);
  }
  __syncthreads();

  for (uint32_t iter = 0;; ++iter) {
    const int slot = (int)(iter & 1u);
    const LocalItem it = sm.item[slot];
    if (it.cnt == 0xFFFFFFFFu) break;
    const uint32_t cnt = it.cnt;
    const uint32_t skew = sm.skew[slot];
    const int hi = it.nbits;                                  // bits [0, hi) are undecided, hi <= 16
    if (tid == PRODUCER) {
      const uint32_t next_i = blockIdx.x + (iter + 1) * gridDim.x;
      stage_item(slot ^ 1, next_i, it_a);
      if (next_i + gridDim.x < num_items) it_b = a.items[next_i + gridDim.x];
    }
    const uint32_t cells = 1u << hi, cmask = cells - 1u;
    K* __restrict__ sk = &sm.stage[slot][0];
    bool sorted = cnt <= 2u * cells + BITMAP_EXTRA;           // block-uniform; otherwise some value certainly occurs too often
    mbar_wait(&sm.bar[slot], (iter >> 1) & 1u);
    const uint32_t aoff = (uint32_t)((reinterpret_cast<uintptr_t>(keys_out + it.off) & 15u) / sizeof(K));
    if (sorted) {
      // ---- phase A: presence bits.  The staged window holds elements [skew, skew + cnt) of 16-byte vectors 0 .. nvec-1
      const uint32_t last = skew + cnt, nvec = (last + E - 1) / E;
      const K hi_bits = (K)(sk[skew] & ~(K)cmask);            // shared by the whole bucket
      uint32_t dupmask = 0;
      // all of a thread's vectors are loaded first, and the atomics of a vector are issued back to back: the returned words
      // are only needed for the duplicate flags, so nothing waits on a shared-memory round trip in between
      uint4 raw[MAXIT];
#pragma unroll
      for (int q = 0; q < MAXIT; ++q) {
        const uint32_t v = q * THREADS + tid;
        raw[q] = v < nvec ? reinterpret_cast<const uint4*>(sk)[v] : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int q = 0; q < MAXIT; ++q) {
        const uint32_t v = q * THREADS + tid;
        if (v < nvec) {
          K e[E];
          memcpy(e, &raw[q], 16);
          uint32_t word[E], bit[E], old[E];
#pragma unroll
          for (int j = 0; j < E; ++j) {
            const uint32_t c = (uint32_t)e[j] & cmask;
            word[j] = c >> 5; bit[j] = 1u << (c & 31u);
          }
          if (v != 0 && v != nvec - 1) {
#pragma unroll
            for (int j = 0; j < E; ++j) old[j] = atomicOr(&sm.bm1[word[j]], bit[j]);
          } else {               // the first / last vector of the window may hold elements outside the bucket
#pragma unroll
            for (int j = 0; j < E; ++j) {
              old[j] = 0;
              if (v * E + j >= skew && v * E + j < last) old[j] = atomicOr(&sm.bm1[word[j]], bit[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < E; ++j) dupmask |= (old[j] & bit[j]) ? (1u << (q * E + j)) : 0u;
        }
      }
      while (dupmask) {          // the few keys whose value was already present: second map, then the short list
        const int b = __ffs(dupmask) - 1;
        dupmask &= dupmask - 1u;
        const uint32_t idx = ((uint32_t)(b / E) * THREADS + tid) * E + (uint32_t)(b % E);
        const uint32_t c = (uint32_t)sk[idx] & cmask;
        const uint32_t bit = 1u << (c & 31u);
        const uint32_t old = atomicOr(&sm.bm2[c >> 5], bit);
        if (old & bit) {
          const uint32_t x = atomicAdd(&sm.nextra, 1u);
          if (x < (uint32_t)BITMAP_EXTRA) sm.extras[x] = c;
        }
      }
      __syncthreads();
      const uint32_t nx = sm.nextra;
      sorted = nx <= (uint32_t)BITMAP_EXTRA;

      // ---- phase B: every thread takes its 32 * WPT consecutive cells out of the maps (and leaves them zeroed)
      uint32_t m1[WPT], m2[WPT];
#pragma unroll
      for (int g = 0; g < WPT / 4; ++g) {
        uint4* p1 = reinterpret_cast<uint4*>(sm.bm1) + tid * (WPT / 4) + g;
        uint4* p2 = reinterpret_cast<uint4*>(sm.bm2) + tid * (WPT / 4) + g;
        const uint4 q1 = *p1, q2 = *p2;
        m1[4 * g] = q1.x; m1[4 * g + 1] = q1.y; m1[4 * g + 2] = q1.z; m1[4 * g + 3] = q1.w;
        m2[4 * g] = q2.x; m2[4 * g + 1] = q2.y; m2[4 * g + 2] = q2.z; m2[4 * g + 3] = q2.w;
        *p1 = make_uint4(0, 0, 0, 0); *p2 = make_uint4(0, 0, 0, 0);
      }
      if (sorted) {
        const uint32_t v_lo = tid * (32u * WPT);
        uint32_t total = 0;
#pragma unroll
        for (int g = 0; g < WPT; ++g) total += __popc(m1[g]) + __popc(m2[g]);
        uint32_t mine_x = 0;         // third and later copies among this thread's cells (almost always none)
        for (uint32_t i = 0; i < nx; ++i) mine_x += (sm.extras[i] - v_lo < 32u * WPT) ? 1u : 0u;
        total += mine_x;
        uint32_t inc = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= (unsigned)o) inc += t;
        }
        if (lane == 31) sm.wt[w] = inc;
        __syncthreads();
        const uint32_t wv = lane < (unsigned)NWARPS ? sm.wt[lane] : 0u;
        uint32_t wi = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
          if (lane >= (unsigned)o) wi += t;
        }
        // ---- phase C: emit this thread's values, highest first, into [.., pos)
        uint32_t pos = aoff + __shfl_sync(0xffffffffu, wi - wv, w) + inc;      // one past this thread's last slot
        if (mine_x == 0) {
#pragma unroll
          for (int g = WPT - 1; g >= 0; --g) {
            uint32_t m = m1[g];
            const uint32_t d = m2[g];
            const K kbase = (K)(hi_bits | (K)(v_lo + 32u * g));
            while (m) {
              const int b = 31 - __clz(m);
              const uint32_t bm = 1u << b;
              m ^= bm;
              const K key = (K)(kbase | (K)b);
              sk[--pos] = key;
              if (d & bm) sk[--pos] = key;
            }
          }
        } else {
#pragma unroll
          for (int g = WPT - 1; g >= 0; --g) {
            uint32_t m = m1[g];
            const uint32_t d = m2[g];
            const uint32_t base = v_lo + 32u * g;
            while (m) {
              const int b = 31 - __clz(m);
              const uint32_t bm = 1u << b;
              m ^= bm;
              const K key = (K)(hi_bits | (K)(base + b));
              if (d & bm) {
                for (uint32_t i = 0; i < nx; ++i)
                  if (sm.extras[i] == base + b) sk[--pos] = key;
                sk[--pos] = key;
              }
              sk[--pos] = key;
            }
          }
        }
      }
      __syncthreads();
      if (tid == 0) sm.nextra = 0;
    }
    if (!sorted && tid == 0) {
      hand_back(a, it);
    }

    // ---- phase D: the bucket sits at sk[aoff ..): shared-memory vector v and output vector v cover the same elements
    if (sorted) {
      K* __restrict__ gdst = keys_out + it.off - aoff;
      const uint32_t total = aoff + cnt, nv = (total + E - 1) / E;
      const bool two = a.tw_out != 0;
      for (uint32_t v = tid; v < nv; v += THREADS) {
        uint4 q = reinterpret_cast<const uint4*>(sk)[v];
        K* e = reinterpret_cast<K*>(&q);
        if (two) {
#pragma unroll
          for (int i = 0; i < E; ++i) e[i] = twiddle_out<K>(e[i], a.tw);
        }
        if (v * E >= aoff && v * E + E <= total) reinterpret_cast<uint4*>(gdst)[v] = q;
        else {
#pragma unroll
          for (int i = 0; i < E; ++i)
            if (v * E + i >= aoff && v * E + i < total) gdst[v * E + i] = e[i];
        }
      }
    }
    __syncthreads();
    if (tid == PRODUCER) it_a = it_b;
  }
}

}  // namespace b200
