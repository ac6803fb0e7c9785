// local_sort.cuh -- small-bucket local sort: a bucket that fits one CTA's shared memory is finished on chip in a
// single read + write sweep.  Replaces do_locrec_radix_sort_keys (msb/src/sort/cuda_radix_sort.h:1332-1620) and, for
// stable single-tile sorts, DeviceRadixSortSingleTileKernel (lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:209-305).
//
// Persistent CTAs take the work items round-robin (items are independent); while one bucket is being sorted the
// next one is already being staged into the other shared-memory slot by a TMA bulk copy.  The result goes to the
// final buffer (in place when the bucket already lives there -- safe because the whole bucket is on chip before
// anything is stored).  Two algorithms, one kernel instantiation each so that neither pays for the other's registers:
//
//  * ALGO_LSD (buckets with <= 16 bits left): LSD passes of 8 bits in shared memory, like the reference
//    (cuda_radix_sort.h:1400-1481,1600-1604): the first pass of an unstable sort ranks with one shared-memory atomicAdd
//    per key, every other pass ranks stably with the match-mask scheme of the stable scatter (scatter.cuh).
//  * ALGO_COUNT (more bits left, e.g. 48 of a 64-bit key): one-shot counting sort.  The top ~log2(count) of the
//    remaining bits index 4-bit shared-memory counters (about one key per cell); a key's atomicAdd returns its rank
//    inside its cell, a prefix sum over the counter words gives every cell's start, and the key goes straight to
//    start + rank: ONE ranking step however many bits remain.  The few keys that share a cell (cells hold <= 15 keys)
//    are ordered by direct comparison inside the cell -- on (key, input index) for stable sorts, which yields exactly
//    the stable order.  A bucket with a cell of more than 15 keys (heavy duplicates / low-entropy bits) is handed to
//    an overflow list that a second ALGO_LSD launch finishes.
#pragma once
#include "async.cuh"
#include "common.cuh"

namespace b200 {

enum { ALGO_LSD = 0, ALGO_COUNT = 1 };

struct LocalArgs {
  void* keys[3]; void* vals[3];          // the ping-pong buffers (LocalItem::src indexes them)
  void* keys_final; void* vals_final;
  const LocalItem* items; const uint32_t* num_items_ptr;
  uint32_t max_items;                    // capacity of every work list (a count above it means the list overflowed: the error flag is up)
  LocalItem* overflow; uint32_t* num_overflow_ptr;     // ALGO_COUNT / bitmap sort: buckets they could not take
  LocalItem* overflow_small; uint32_t* num_overflow_small_ptr; uint32_t overflow_small_cap;   // ... those of at most overflow_small_cap keys (nullptr: all to `overflow`)
  uint32_t* error_ptr;                   // MsbCounters::error
  LocalItem* dense; uint32_t* num_dense_ptr;           // rank sort: buckets passed on to its dense variant (nullptr: none)
  int tw_in;                             // keys still in caller form (single-tile sorts)
  int tw_out;
  int begin_bit;                         // lowest bit to sort (0 for MSB items)
  Twiddle tw;
};

constexpr int COUNT_MAX_BITS = 16;       // cells when <= 16 bits remain: one cell per remaining key value (no in-cell ordering needed)
constexpr int COUNT_FIT_BITS = 13;       // cells when more bits remain: ~log2(count), at most 8192 (about one key per cell)

// MAXBITS = 16 only where the exact-cell mode is used (4-byte keys-only buckets with <= 16 bits left): 48 KB of counters; every
// other instantiation keeps the compact 13-bit table (6 KB), which leaves the L1 / shared-memory split alone.
template <int THREADS, int MAXBITS>
struct CountSmem {
  static constexpr int WORDS = (1 << MAXBITS) / 8;
  static constexpr int NSEG = (WORDS / 4 + THREADS - 1) / THREADS;   // consecutive 4-word groups a thread owns in the prefix scan
  alignas(16) uint32_t nib[WORDS];     // 8 counters of 4 bits per word
  alignas(16) uint16_t wpre[WORDS];    // number of keys in all earlier words
  uint32_t wt[32];                     // per-warp totals of the scan
};
template <typename K, int VB> struct CountBits { static constexpr int value = (sizeof(K) == 4 && VB == 0) ? COUNT_MAX_BITS : COUNT_FIT_BITS; };

template <int THREADS>
struct LsdSmem {
  static constexpr int WARPS = THREADS / 32;
  alignas(16) uint32_t match[2][WARPS * RADIX];   // per-warp match masks, two alternating sets (stable ranking)
  alignas(16) uint16_t wcnt[WARPS * RADIX];       // per-warp counters, later per-warp start positions
  uint32_t cnt[RADIX];                            // block counters of the unordered first pass
  uint32_t bin_start[RADIX];
  uint32_t scratch[8];
  uint32_t red[4];                                // OR (low, high word) and AND (low, high word) of the bucket's keys
};

template <typename K, int VB, int THREADS, int IPT, int ALGO>
struct LocalSmem {
  static constexpr int CAP = THREADS * IPT;
  static constexpr int SLACK = 16 / sizeof(K);
  using V = typename ValType<VB>::type;
  static constexpr int VSLACK = 16 / sizeof(V);
  alignas(16) K stage[2][CAP + SLACK];
  alignas(16) V vstage[VB ? 2 : 1][VB ? CAP + VSLACK : 1];
  alignas(16) uint16_t origin[ALGO == ALGO_COUNT ? CAP : 8];    // stable counting sort: input index of the key at each position
  typename std::conditional<ALGO == ALGO_LSD, LsdSmem<THREADS>, CountSmem<THREADS, CountBits<K, VB>::value>>::type rank;
  alignas(8) uint64_t bar[2];
  LocalItem item[2];
  uint32_t skew[2], vskew[2];
};

// ---------------------------------------------------------------------------------------------------------------
// ALGO_COUNT.  Returns false (nothing modified) if a cell overflowed; on success the sorted bucket is at sk[0..cnt),
// sv[0..cnt).
// ---------------------------------------------------------------------------------------------------------------
// Exact-cell keys-only buckets land at sk[aoff .. aoff+cnt) (aoff = the bucket's element offset inside its 16-byte granule of the
// output), so that the write-out can move aligned 16-byte vectors from shared memory to the output.
template <typename K, int VB, int THREADS, int ROWS, bool STABLE>
__device__ __forceinline__ bool count_sort_item(K* __restrict__ sk, typename ValType<VB>::type* __restrict__ sv, uint16_t* __restrict__ origin,
                                                uint32_t skew, uint32_t vskew, uint32_t cnt, int lo, int hi, bool tw_in, const Twiddle& tw,
                                                CountSmem<THREADS, CountBits<K, VB>::value>& cs, uint32_t aoff) {
  using V = typename ValType<VB>::type;
  constexpr int NWARPS = THREADS / 32;
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const int rows = (int)((cnt + THREADS - 1) / THREADS);      // block-uniform: rows past the bucket are skipped without per-thread tests
  constexpr int MAXBITS = CountBits<K, VB>::value;
  constexpr int NSEG = CountSmem<THREADS, MAXBITS>::NSEG;
  int cbits = 32 - __clz(cnt);
  cbits = cbits < 6 ? 6 : (cbits > COUNT_FIT_BITS ? COUNT_FIT_BITS : cbits);
  if (hi - lo <= MAXBITS) cbits = hi - lo;                 // few bits left: one cell per key value
  const int vshift = hi - cbits;
  const uint32_t vmask = (1u << cbits) - 1u;
  const uint32_t words = cbits > 5 ? 1u << (cbits - 3) : 4u;
  const uint32_t groups = words / 4;

  for (uint32_t i = tid; i < groups; i += THREADS) reinterpret_cast<uint4*>(cs.nib)[i] = make_uint4(0, 0, 0, 0);
  K key[ROWS]; uint32_t pos[ROWS]; V val[VB ? ROWS : 1];
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    const uint32_t idx = j * THREADS + tid;
    K k = (K)0;
    if (j < rows && idx < cnt) { k = sk[skew + idx]; if (VB) val[j] = sv[vskew + idx]; }
    key[j] = k;
  }
  if (tw_in) {
#pragma unroll
    for (int j = 0; j < ROWS; ++j) key[j] = twiddle_in<K>(key[j], tw);
  }
  __syncthreads();
  int ovf = 0;
#pragma unroll
  for (int j = 0; j < ROWS; ++j)
    if (j < rows && (uint32_t)(j * THREADS) + tid < cnt) {
      const uint32_t v = (uint32_t)(key[j] >> vshift) & vmask;
      const uint32_t sh = (v & 7u) * 4u;
      const uint32_t old = atomicAdd(&cs.nib[v >> 3], 1u << sh);
      pos[j] = (old >> sh) & 15u;
      ovf |= (pos[j] == 15u);
    }
  if (__syncthreads_or(ovf)) return false;

  // ---- exclusive prefix over the counter words; thread t owns the NSEG consecutive 4-word groups t*NSEG ..
  {
    // per-word nibble sums (a word holds <= 8 * 15 keys; a 4-word group can exceed 255, so the words are summed separately)
    // acc + the eight nibbles of x: two byte-wise dot products with ones (IDP.4A accumulates for free; measured 2.4 % of a
    // 2^28-key sort against the and/shift/multiply form)
    auto nadd = [](uint32_t x, uint32_t acc) { return __dp4a(x & 0x0F0F0F0Fu, 0x01010101u, __dp4a((x >> 4) & 0x0F0F0F0Fu, 0x01010101u, acc)); };
    uint32_t tsum = 0;
#pragma unroll
    for (int g = 0; g < NSEG; ++g) {
      const uint32_t gi = tid * NSEG + g;
      if (gi < groups) {
        const uint4 q = reinterpret_cast<const uint4*>(cs.nib)[gi];
        tsum = nadd(q.w, nadd(q.z, nadd(q.y, nadd(q.x, tsum))));
      }
    }
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) cs.wt[w] = inc;
    __syncthreads();
    const uint32_t wv = lane < (unsigned)NWARPS ? cs.wt[lane] : 0u;
    uint32_t wi = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= (unsigned)o) wi += t;
    }
    uint32_t run = __shfl_sync(0xffffffffu, wi - wv, w) + inc - tsum;
#pragma unroll
    for (int g = 0; g < NSEG; ++g) {
      const uint32_t gi = tid * NSEG + g;
      if (gi < groups) {
        const uint4 q = reinterpret_cast<const uint4*>(cs.nib)[gi];
        const uint32_t p0 = run; run = nadd(q.x, run);
        const uint32_t p1 = run; run = nadd(q.y, run);
        const uint32_t p2 = run; run = nadd(q.z, run);
        const uint32_t p3 = run; run = nadd(q.w, run);
        reinterpret_cast<uint2*>(cs.wpre)[gi] = make_uint2(p0 | (p1 << 16), p2 | (p3 << 16));
      }
    }
  }
  __syncthreads();

  // ---- scatter to cell start + rank in cell
  uint32_t cell0[ROWS];          // packed: start of the key's cell | keys in the cell << 16 (only where cells hold distinct keys)
  const bool order_cells = STABLE || hi - lo > cbits;
#if B200_LOCAL_LEAN_EXACT
  if (!order_cells) {            // exact cells: nothing else to record
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const uint32_t idx = j * THREADS + tid;
      if (j < rows && idx < cnt) {
        const uint32_t v = (uint32_t)(key[j] >> vshift) & vmask;
        const uint32_t lowc = cs.nib[v >> 3] & ((1u << ((v & 7u) * 4u)) - 1u);
        const uint32_t c0 = __dp4a(lowc & 0x0F0F0F0Fu, 0x01010101u, __dp4a((lowc >> 4) & 0x0F0F0F0Fu, 0x01010101u, (uint32_t)cs.wpre[v >> 3]));
        const uint32_t q = pos[j] + c0;
        sk[aoff + q] = key[j];
        if (VB) sv[q] = val[j];      // every thread read its values at the top (a barrier ago): the slot is free
      }
    }
    __syncthreads();
    return true;
  }
#endif
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    const uint32_t idx = j * THREADS + tid;
    if (j < rows && idx < cnt) {
      const uint32_t v = (uint32_t)(key[j] >> vshift) & vmask;
      const uint32_t sh = (v & 7u) * 4u;
      const uint32_t wd = cs.nib[v >> 3];
      const uint32_t lowc = wd & ((1u << sh) - 1u);
      const uint32_t c0 = __dp4a(lowc & 0x0F0F0F0Fu, 0x01010101u, __dp4a((lowc >> 4) & 0x0F0F0F0Fu, 0x01010101u, (uint32_t)cs.wpre[v >> 3]));
      cell0[j] = c0 | (((wd >> sh) & 15u) << 16);
      pos[j] += c0;
      sk[aoff + pos[j]] = key[j];
      if (STABLE) origin[pos[j]] = (uint16_t)idx;
    }
  }

  // ---- keys that share a cell: order them by direct comparison on bits [lo, hi) (and the input index when stable)
  if (order_cells) {
    __syncthreads();
    uint32_t npos[ROWS];
    const int up = (int)sizeof(K) * 8 - hi, dn = up + lo;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      npos[j] = 0xFFFFFFFFu;
      const uint32_t idx = j * THREADS + tid;
      if (j < rows && idx < cnt) {
        const uint32_t c = cell0[j] >> 16, c0 = cell0[j] & 0xFFFFu;
        if (c > 1) {
          const uint32_t r = pos[j] - c0;
          uint32_t t = 0;
          const K mine = (K)((K)(key[j] << up) >> dn);
          for (uint32_t i = 0; i < c; ++i) {
            const K o = (K)((K)(sk[c0 + i] << up) >> dn);
            if (STABLE) t += (o < mine || (o == mine && (uint32_t)origin[c0 + i] < idx)) ? 1u : 0u;
            else t += (o < mine || (o == mine && i < r)) ? 1u : 0u;
          }
          if (t != r) { npos[j] = c0 + t; pos[j] = c0 + t; }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
      if (npos[j] != 0xFFFFFFFFu) sk[npos[j]] = key[j];
  }
  if (VB) {
    // every thread read its values at the top: the value slot is free to receive the sorted order
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
      if (j < rows && (uint32_t)(j * THREADS) + tid < cnt) sv[pos[j]] = val[j];
  }
  __syncthreads();
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// ALGO_LSD.  On return the sorted bucket is at sk[0..cnt), sv[0..cnt).
// ---------------------------------------------------------------------------------------------------------------
// SKEW (compile time): the bucket was handed back by a one-shot kernel, i.e. its digits are skewed (heavy duplicates, low-entropy
// bits).  Same-address shared-memory atomics serialise (measured 8-10 wavefronts per instruction on such buckets,
// profiles/r02_lsd_entropy2.txt), so the lanes of a warp that hold the row's HOT digit are found with one vote and share one
// atomic / one counter update; the other lanes rank as usual.  Uniform digits are better off without the extra votes (measured:
// +25 % on the merged runs of a 2^24-key sort), hence two instantiations.
template <typename K, int VB, int THREADS, int ROWS, bool STABLE, bool SKEW = false>
__device__ __forceinline__ void lsd_sort_item(K* __restrict__ sk, typename ValType<VB>::type* __restrict__ sv, uint32_t skew, uint32_t vskew,
                                              uint32_t cnt, int lo, int hi, bool tw_in, const Twiddle& tw, LsdSmem<THREADS>& sm) {
  using V = typename ValType<VB>::type;
  constexpr int WARPS = THREADS / 32;
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const int rows = (int)((cnt + THREADS - 1) / THREADS);
  const int passes = hi > lo ? (hi - lo + 7) / 8 : 0;
  const uint32_t obase = w * (uint32_t)rows * 32u + lane;          // warp-contiguous layout of the stable passes
  K key[ROWS]; V val[VB ? ROWS : 1]; uint32_t pos[ROWS];
  const bool first_ordered = STABLE || passes == 0;
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    K k = (K)~(K)0;
    if (j < rows) {
      const uint32_t idx = first_ordered ? obase + j * 32u : (uint32_t)j * THREADS + tid;
      if (idx < cnt) { k = sk[skew + idx]; if (VB) val[j] = sv[vskew + idx]; }
    }
    key[j] = k;
  }
  if (tw_in) {       // padding keys must stay all ones
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const uint32_t idx = first_ordered ? obase + j * 32u : (uint32_t)j * THREADS + tid;
      if (j < rows && idx < cnt) key[j] = twiddle_in<K>(key[j], tw);
    }
  }
  // Constant-digit short circuit (cf. CUB's short_circuit, lsb/cub/cub/agent/agent_radix_sort_downsweep.cuh:701-724): a pass whose digit
  // is the same for every key of the bucket changes nothing and is skipped.  Buckets of heavily duplicated inputs (one hot value, or
  // a handful of values that differ in a few bits) otherwise pay a full ranking pass for every 8 bits that remain.
  uint32_t active = 0;                                   // bit p: pass p (bits [lo + 8p, lo + 8p + 8)) has work to do
  if (passes > 0) {
    K kor = (K)0, kand = (K)~(K)0;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const uint32_t idx = first_ordered ? obase + j * 32u : (uint32_t)j * THREADS + tid;
      if (j < rows && idx < cnt) { kor |= key[j]; kand &= key[j]; }
    }
    if (tid == 0) { sm.red[0] = 0u; sm.red[1] = 0u; sm.red[2] = 0xFFFFFFFFu; sm.red[3] = 0xFFFFFFFFu; }
    __syncthreads();
    const uint32_t o0 = __reduce_or_sync(0xffffffffu, (uint32_t)kor), a0 = __reduce_and_sync(0xffffffffu, (uint32_t)kand);
    uint32_t o1 = 0u, a1 = 0xFFFFFFFFu;
    if (sizeof(K) == 8) {
      o1 = __reduce_or_sync(0xffffffffu, (uint32_t)((unsigned long long)kor >> 32));
      a1 = __reduce_and_sync(0xffffffffu, (uint32_t)((unsigned long long)kand >> 32));
    }
    if (lane == 0) {
      atomicOr(&sm.red[0], o0); atomicAnd(&sm.red[2], a0);
      if (sizeof(K) == 8) { atomicOr(&sm.red[1], o1); atomicAnd(&sm.red[3], a1); }
    }
    __syncthreads();
    const unsigned long long diff = (((unsigned long long)sm.red[1] << 32) | sm.red[0]) ^ (((unsigned long long)sm.red[3] << 32) | sm.red[2]);
    for (int p = 0; p < passes; ++p) {
      const int shift = lo + 8 * p;
      const uint32_t mask = (1u << (hi - shift < 8 ? hi - shift : 8)) - 1u;
      if (((uint32_t)(diff >> shift) & mask) != 0u) active |= 1u << p;
    }
  }
  if (active == 0) {      // nothing to sort: (transformed) copy through the slot
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const uint32_t idx = first_ordered ? obase + j * 32u : (uint32_t)j * THREADS + tid;
      if (j < rows && idx < cnt) { sk[idx] = key[j]; if (VB) sv[idx] = val[j]; }
    }
    __syncthreads();
    return;
  }

  if (!STABLE) {
    // ---- first pass of an unstable sort: one shared-memory atomicAdd per key
    const int p0 = __ffs(active) - 1;
    active &= active - 1u;
    const int shift = lo + 8 * p0;
    const uint32_t mask = (1u << (hi - shift < 8 ? hi - shift : 8)) - 1u;
    if (tid < RADIX) sm.cnt[tid] = 0;
    __syncthreads();
    if (!SKEW) {
#pragma unroll
      for (int j = 0; j < ROWS; ++j)
        if (j < rows && (uint32_t)j * THREADS + tid < cnt) pos[j] = atomicAdd(&sm.cnt[digit_of<K>(key[j], shift, mask)], 1u);
    } else {
      unsigned hot_d = 0;
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int j = 0; j < ROWS; ++j)
        if (j < rows) {      // block-uniform
          const bool v = (uint32_t)j * THREADS + tid < cnt;
          const unsigned d = digit_of<K>(key[j], shift, mask);
          const unsigned hot = __ballot_sync(0xffffffffu, v && d == hot_d);      // one atomic for all lanes on the warp's hot digit
          if (v && d == hot_d) {
            unsigned b = 0;
            if ((hot & lt) == 0u) b = atomicAdd(&sm.cnt[d], (uint32_t)__popc(hot));
            pos[j] = __shfl_sync(hot, b, __ffs(hot) - 1) + __popc(hot & lt);
          } else if (v) {
            pos[j] = atomicAdd(&sm.cnt[d], 1u);
          }
          // a small hot group: try the digit of the row's first valid lane next
          if (__popc(hot) < 8) { const unsigned vm = __ballot_sync(0xffffffffu, v); hot_d = vm ? __shfl_sync(0xffffffffu, d, __ffs(vm) - 1) : hot_d; }
        }
    }
    __syncthreads();
    const uint32_t total = tid < RADIX ? sm.cnt[tid] : 0u;
    const uint32_t excl = block_excl_scan_256(total, sm.scratch);
    if (tid < RADIX) sm.bin_start[tid] = excl;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
      if (j < rows && (uint32_t)j * THREADS + tid < cnt) {
        const uint32_t q = pos[j] + sm.bin_start[digit_of<K>(key[j], shift, mask)];
        sk[q] = key[j];
        if (VB) sv[q] = val[j];
      }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int j = 0; j < ROWS; ++j) {
        K k = (K)~(K)0;
        if (j < rows) {
          const uint32_t idx = obase + j * 32u;
          if (idx < cnt) { k = sk[idx]; if (VB) val[j] = sv[idx]; }
        }
        key[j] = k;
      }
    }
  }
  while (active) {
    // ---- stable pass
    const int p = __ffs(active) - 1;
    active &= active - 1u;
    const int shift = lo + 8 * p;
    const uint32_t mask = (1u << (hi - shift < 8 ? hi - shift : 8)) - 1u;
    // (the match masks are all zero here: zeroed once at kernel start, and every row's leader clears the word it used)
    uint4* zc = reinterpret_cast<uint4*>(sm.wcnt);
    for (int i = tid; i < WARPS * RADIX / 8; i += THREADS) zc[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    uint16_t* wc = sm.wcnt + w * RADIX;
    const unsigned lt = (1u << lane) - 1u, lbit = 1u << lane;
    unsigned hot_d = 0;        // SKEW: the most frequent digit of the warp's previous row
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      if (j < rows) {        // block-uniform
        uint32_t* wm = sm.match[j & 1] + w * RADIX;
        const unsigned d = digit_of<K>(key[j], shift, mask);       // padding keys (all ones) sit last and rank last
        const bool is_hot = SKEW && d == hot_d;
        const unsigned hot = SKEW ? __ballot_sync(0xffffffffu, is_hot) : 0u;
        if (!is_hot) atomicOr(&wm[d], lbit);
        __syncwarp();
        const unsigned peers = is_hot ? hot : wm[d];
        __syncwarp();
        const unsigned below = __popc(peers & lt);
        unsigned b = 0;
        if (below == 0) { b = wc[d]; wc[d] = (uint16_t)(b + __popc(peers)); if (!is_hot) wm[d] = 0; }
        b = __shfl_sync(0xffffffffu, b, __ffs(peers) - 1);
        pos[j] = b + below;
        if (SKEW) hot_d = __reduce_max_sync(0xffffffffu, ((unsigned)__popc(peers) << 8) | d) & 0xFFu;
      }
    }
    __syncthreads();
    uint32_t my_total = 0;
    if (tid < RADIX) {
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) my_total += sm.wcnt[ww * RADIX + tid];
      if (tid == mask) my_total -= (uint32_t)rows * THREADS - cnt;     // padding
    }
    const uint32_t excl = block_excl_scan_256(my_total, sm.scratch);
    if (tid < RADIX) {
      uint32_t run = excl;
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) {
        const uint32_t c = sm.wcnt[ww * RADIX + tid];
        sm.wcnt[ww * RADIX + tid] = (uint16_t)run;
        run += c;
      }
    }
    __syncthreads();
    // every thread has its keys in registers: the slot is the exchange buffer (positions 0..cnt-1)
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
      if (j < rows) {
        const uint32_t q = pos[j] + wc[digit_of<K>(key[j], shift, mask)];
        if (q < cnt) { sk[q] = key[j]; if (VB) sv[q] = val[j]; }     // padding ranks after every real key
      }
    __syncthreads();
    if (active) {             // read back in warp-contiguous order for the next pass
#pragma unroll
      for (int j = 0; j < ROWS; ++j) {
        K k = (K)~(K)0;
        if (j < rows) {
          const uint32_t idx = obase + j * 32u;
          if (idx < cnt) { k = sk[idx]; if (VB) val[j] = sv[idx]; }
        }
        key[j] = k;
      }
      // no barrier needed here: the ranking synchronises before anybody scatters again
    }
  }
}

#ifndef B200_LOCAL_OCC384
#define B200_LOCAL_OCC384 2
#endif
#ifndef B200_LOCAL_LEAN_EXACT
#define B200_LOCAL_LEAN_EXACT 1
#endif
#ifndef B200_LOCAL_VEC_OUT
#define B200_LOCAL_VEC_OUT 1
#endif
// A bucket the one-shot kernels could not take (heavy duplicates, skewed bits) goes to the LSD kernels: by size, because a small bucket
// in the large configuration pays that configuration's fixed cost per pass (24 warps' counters, four block barriers) for a handful of rows.
__device__ __forceinline__ void hand_back(const LocalArgs& a, const LocalItem& it) {
  const bool small = a.overflow_small != nullptr && it.cnt <= a.overflow_small_cap;
  const uint32_t o = atomicAdd(small ? a.num_overflow_small_ptr : a.num_overflow_ptr, 1u);
  if (o < a.max_items) (small ? a.overflow_small : a.overflow)[o] = it; else atomicOr(a.error_ptr, 2u);
}

// SKEW: the launch serves the handed-back buckets, whose digits are known to be skewed (see lsd_sort_item).
template <typename K, int VB, int THREADS, int IPT, int ALGO, bool STABLE, bool SKEW = false>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? (IPT > 6 ? 3 : 4) : THREADS == 384 ? ((VB == 0 && ALGO == ALGO_LSD) ? 3 : B200_LOCAL_OCC384) : (sizeof(LocalSmem<K, VB, THREADS, IPT, ALGO>) <= 113 * 1024 && THREADS <= 768) ? 2 : 1)) local_sort_kernel(const __grid_constant__ LocalArgs a) {
  pdl_wait();
  using V = typename ValType<VB>::type;
  using SM = LocalSmem<K, VB, THREADS, IPT, ALGO>;
  constexpr unsigned PRODUCER = THREADS - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x;
  const uint32_t num_items = min(*a.num_items_ptr, a.max_items);
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

  if (ALGO == ALGO_LSD) {
    LsdSmem<THREADS>& ls = *reinterpret_cast<LsdSmem<THREADS>*>(&sm.rank);
    uint4* z = reinterpret_cast<uint4*>(ls.match);
    for (int i = tid; i < 2 * (THREADS / 32) * RADIX / 4; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  // the producer knows its next item one iteration ahead, so the prefetch into the free slot goes out at the top
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
    const uint32_t skew = sm.skew[slot], vskew = VB ? sm.vskew[slot] : 0;
    const int lo = a.begin_bit, hi = it.nbits;
    if (tid == PRODUCER) {
      const uint32_t next_i = blockIdx.x + (iter + 1) * gridDim.x;
      stage_item(slot ^ 1, next_i, it_a);          // the other slot is free: its bucket was written out last iteration
      if (next_i + gridDim.x < num_items) it_b = a.items[next_i + gridDim.x];
    }
    mbar_wait(&sm.bar[slot], (iter >> 1) & 1u);
    K* __restrict__ sk = &sm.stage[slot][0];
    V* __restrict__ sv = &sm.vstage[VB ? slot : 0][0];

    bool sorted = true;
    bool vec = false; uint32_t aoff = 0;
    if (ALGO == ALGO_COUNT) {
      vec = B200_LOCAL_VEC_OUT && !STABLE && VB == 0 && hi - lo <= CountBits<K, VB>::value;       // exact cells: no in-cell ordering step touches sk
      aoff = vec ? (uint32_t)((reinterpret_cast<uintptr_t>(keys_out + it.off) & 15u) / sizeof(K)) : 0u;
      sorted = count_sort_item<K, VB, THREADS, IPT, STABLE>(sk, sv, sm.origin, skew, vskew, cnt, lo, hi, a.tw_in != 0, a.tw,
                                                             *reinterpret_cast<CountSmem<THREADS, CountBits<K, VB>::value>*>(&sm.rank), aoff);
      if (!sorted && tid == 0) hand_back(a, it);
    } else {
      lsd_sort_item<K, VB, THREADS, IPT, STABLE, SKEW>(sk, sv, skew, vskew, cnt, lo, hi, a.tw_in != 0, a.tw, *reinterpret_cast<LsdSmem<THREADS>*>(&sm.rank));
    }

    // ---- coalesced write-out of the sorted bucket
    if (sorted && vec) {
      // the bucket sits at sk[aoff ..): shared-memory vector v and output vector v cover the same elements, both 16-byte aligned
      constexpr uint32_t E = 16 / sizeof(K);
      K* __restrict__ gdst = keys_out + it.off - aoff;
      const uint32_t total = aoff + cnt, nvec = (total + E - 1) / E;
      const bool two = a.tw_out != 0;
      for (uint32_t v = tid; v < nvec; v += THREADS) {
        uint4 q = reinterpret_cast<const uint4*>(sk)[v];
        K* e = reinterpret_cast<K*>(&q);
        if (two) {
#pragma unroll
          for (uint32_t i = 0; i < E; ++i) e[i] = twiddle_out<K>(e[i], a.tw);
        }
        if (v * E >= aoff && v * E + E <= total) reinterpret_cast<uint4*>(gdst)[v] = q;
        else {
#pragma unroll
          for (uint32_t i = 0; i < E; ++i)
            if (v * E + i >= aoff && v * E + i < total) gdst[v * E + i] = e[i];
        }
      }
    } else if (sorted) {
      if (a.tw_out) {
        for (uint32_t pidx = tid; pidx < cnt; pidx += THREADS) {
          keys_out[it.off + pidx] = twiddle_out<K>(sk[pidx], a.tw);
          if (VB) vals_out[it.off + pidx] = sv[pidx];
        }
      } else {
        for (uint32_t pidx = tid; pidx < cnt; pidx += THREADS) {
          keys_out[it.off + pidx] = sk[pidx];
          if (VB) vals_out[it.off + pidx] = sv[pidx];
        }
      }
    }
    __syncthreads();
    if (tid == PRODUCER) it_a = it_b;
  }
}

}  // namespace b200
