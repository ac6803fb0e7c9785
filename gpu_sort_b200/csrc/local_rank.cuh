// local_rank.cuh -- on-chip finish of buckets with at most 16 undecided bits: one-shot rank by presence bits.
// Replaces do_locrec_radix_sort_keys (msb/src/sort/cuda_radix_sort.h:1332-1620: an 8-bit counting sort followed by
// cub::BlockRadixSort passes) for 4-byte keys, keys-only and pairs, unstable (MSB API) and stable (DeviceRadixSort API).
//
// A bucket of ~4096 keys with 16 bits left touches ~6 % of its 65 536 possible cells, so almost every key is alone in
// its cell and its final position is simply the number of occupied cells below it:
//   A  every key sets its cell's bit with one shared-memory atomicOr (lo[cell / 16], bits 0-15); the few keys that find
//      the bit already set (3 % for uniform keys) set the "second copy" bit (bits 16-31); third and later copies go
//      to a short list (more than RANK_EXTRA of them: the bucket is handed to the LSD kernel through the overflow list);
//   B  a block-wide exclusive scan of popc(lo[]) gives hi[e] = keys in all earlier 16-cell entries;
//   C  rank(key) = hi[e] + popc(lo[e] & bits below the key's cell) -- ONE step however many bits remain, uniform
//      control flow, every key of a thread independent of the others.  Keys sharing a cell: unstable sorts place them in
//      arrival order; stable sorts let the copies meet in a per-slot table and order them by input index;
//   D  coalesced 16-byte stores from the consumed staging slot (laid out so shared-memory and output vectors coincide).
// The kernel is bound by shared-memory wavefronts (four random 4-byte accesses per key: atomicOr, lo, hi, reorder
// store); see DESIGN.md "Kernels" for the measured numbers.
#pragma once
#include "async.cuh"
#include "common.cuh"
#include "local_sort.cuh"

namespace b200 {

constexpr int RANK_BITS = 16;
constexpr int RANK_ENT = (1 << RANK_BITS) / 16;       // 16 cells per entry
constexpr int RANK_DENSE_BITS = 15;    // DENSE: 2^15 cells x 4-bit counters fill lo[]
constexpr int RANK_EXTRA = 15;         // (<= 15: the scan counts the listed copies of an entry in a 4-bit field)

template <typename K, int VB, int THREADS, int IPT, bool STABLE>
struct RankSmem {
  static constexpr int CAP = THREADS * IPT;
  using V = typename ValType<VB>::type;
  static constexpr int EK = 16 / sizeof(K), EV = 16 / sizeof(V);
  static constexpr int EPT = RANK_ENT / THREADS;        // entries a thread owns in the scan
  static_assert(RANK_ENT % THREADS == 0 && EPT % 4 == 0, "a thread owns whole 16-byte vectors of lo[]");
  alignas(16) K stage[2][CAP + 2 * EK];
  alignas(16) V vstage[VB ? 2 : 1][VB ? CAP + 2 * EV : 1];
  alignas(16) uint32_t lo[RANK_ENT];          // bits 0-15: cell occupied; bits 16-31: cell holds a second copy
  alignas(16) uint16_t hi[RANK_ENT];          // keys in earlier entries; bit 15: the entry holds third-or-later copies
  alignas(16) uint16_t origin[STABLE ? CAP : 8];          // stable sorts: input index of every copy of a shared cell, by slot
  static_assert(CAP < 32768, "hi[] keeps a flag in bit 15");
  uint32_t extras[RANK_EXTRA];                // third and later copies: cell | input index << 16, in arrival order
  uint32_t nextra;
  uint32_t wt[32];
  alignas(8) uint64_t bar[2];
  LocalItem item[2];
  uint32_t skew[2], vskew[2];
};

// A copy of a block-uniform value the compiler cannot see through: the row tests of each phase (`j < rows_full`) are then
// evaluated where they are used (one compare per row) instead of being computed once, packed into a register bit mask and
// unpacked again in every phase (measured: 6 instructions per key, profiles/r02_rank_v2.txt).
// sum of the eight 4-bit fields of x
__device__ __forceinline__ uint32_t nibble_sum(uint32_t x) { return __dp4a(x & 0x0F0F0F0Fu, 0x01010101u, __dp4a((x >> 4) & 0x0F0F0F0Fu, 0x01010101u, 0u)); }

__device__ __forceinline__ uint32_t opaque(uint32_t x) { uint32_t y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <typename K, int VB, int THREADS, int IPT, int OCC, bool STABLE, bool DENSE = false>
__global__ void __launch_bounds__(THREADS, OCC) rank_sort_kernel(const __grid_constant__ LocalArgs a) {
  pdl_wait();
  using SM = RankSmem<K, VB, THREADS, IPT, STABLE>;
  using V = typename SM::V;
  constexpr unsigned PRODUCER = THREADS - 1;
  constexpr int EK = SM::EK, EV = SM::EV, EPT = SM::EPT, NWARPS = THREADS / 32;
  constexpr bool ORDER = STABLE;                  // keys sharing a cell keep their input order
  static_assert(IPT <= 32, "one flag bit per key of a thread");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const uint32_t num_items = min(*a.num_items_ptr, a.max_items);
  K* __restrict__ keys_out = reinterpret_cast<K*>(a.keys_final);
  V* __restrict__ vals_out = reinterpret_cast<V*>(a.vals_final);

  auto stage_item = [&](int slot, uint32_t i, const LocalItem& it) {
    if (i < num_items) {
      if (!DENSE && a.dense != nullptr) {
        // more than one key per twelve cells: third copies of a value are no longer rare (measured: at one per eight most buckets
        // exceed the short list) -- the 4-bit-counter variant takes the bucket (passed on before anything of it is read)
        const int nb = (int)it.nbits - a.begin_bit;
        if (nb <= RANK_DENSE_BITS && (unsigned long long)it.cnt * 12ull > (1ull << (nb > 0 ? nb : 0))) {
          const uint32_t o = atomicAdd(a.num_dense_ptr, 1u);
          if (o < a.max_items) a.dense[o] = it; else atomicOr(a.error_ptr, 2u);
          LocalItem skip = it; skip.cnt = 0xFFFFFFFEu;      // "nothing staged": the block moves on to its next bucket
          sm.item[slot] = skip;
          return;
        }
      }
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
      LocalItem none{}; none.cnt = 0xFFFFFFFFu;
      sm.item[slot] = none;
    }
  };

  for (int i = tid; i < RANK_ENT / 4; i += THREADS) reinterpret_cast<uint4*>(sm.lo)[i] = make_uint4(0, 0, 0, 0);
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

  uint32_t phase = 0u;                       // bit s: mbarrier phase parity of staging slot s (a slot whose bucket was passed on is not armed)
  for (uint32_t iter = 0;; ++iter) {
    const int slot = (int)(iter & 1u);
    const LocalItem it = sm.item[slot];
    if (it.cnt == 0xFFFFFFFFu) break;
    const uint32_t cnt = it.cnt;
    const uint32_t skew = sm.skew[slot], vskew = VB ? sm.vskew[slot] : 0;
    const int lo_bit = a.begin_bit, nb = (int)it.nbits - lo_bit;      // bits [lo_bit, lo_bit + nb) are undecided
    if (tid == PRODUCER) {
      const uint32_t next_i = blockIdx.x + (iter + 1) * gridDim.x;
      stage_item(slot ^ 1, next_i, it_a);
      if (next_i + gridDim.x < num_items) it_b = a.items[next_i + gridDim.x];
    }
    K* __restrict__ sk = &sm.stage[slot][0];
    V* __restrict__ sv = &sm.vstage[VB ? slot : 0][0];
    const uint32_t cells = nb > 0 ? 1u << (nb > RANK_BITS ? RANK_BITS : nb) : 1u, cmask = cells - 1u;
    // entries of lo[] in use: 16 cells per entry (presence + second-copy bits), DENSE: 8 cells per entry (4-bit counters)
    const uint32_t nent = DENSE ? (cells > 8u ? cells >> 3 : 1u) : (cells > 16u ? cells >> 4 : 1u);
    bool sorted = DENSE ? (nb <= RANK_DENSE_BITS && cnt <= 15u * cells) : (nb <= RANK_BITS && cnt <= 2u * cells + RANK_EXTRA);      // block-uniform
    if (cnt == 0xFFFFFFFEu) {                  // passed on to the dense variant by the producer: nothing was staged
      __syncthreads();
      if (tid == PRODUCER) it_a = it_b;
      continue;
    }
    mbar_wait(&sm.bar[slot], (phase >> slot) & 1u);
    phase ^= 1u << slot;
    // the sorted bucket is built at sk[aoff ..) / sv[voff ..): same 16-byte phase as its place in the output arrays
    const uint32_t aoff = (uint32_t)((reinterpret_cast<uintptr_t>(keys_out + it.off) & 15u) / sizeof(K));
    const uint32_t voff = VB ? (uint32_t)((reinterpret_cast<uintptr_t>(vals_out + it.off) & 15u) / sizeof(V)) : 0u;

    if (sorted) {
      // ---- A: keys (and values) to registers, presence bits.  Row j of the bucket = keys j * THREADS .. : rows below
      // rows_full are complete (no per-key tests anywhere), row rows_full holds the last `rem` keys.
      const uint32_t rows_full = cnt / THREADS, rem = cnt - rows_full * THREADS;
      const bool has_last = tid < rem;                       // this thread holds a key of the incomplete row
      K key[IPT], key_last = (K)0; V val[VB ? IPT : 1], val_last = (V)0;
#pragma unroll
      for (int j = 0; j < IPT; ++j) {                        // (rows past the bucket read stale staging data that is never used)
        key[j] = sk[skew + j * THREADS + tid];
        if (VB) val[j] = sv[vskew + j * THREADS + tid];
      }
      if (has_last) { key_last = sk[skew + rows_full * THREADS + tid]; if (VB) val_last = sv[vskew + rows_full * THREADS + tid]; }
      if (a.tw_in) {
#pragma unroll
        for (int j = 0; j < IPT; ++j) key[j] = twiddle_in<K>(key[j], a.tw);
        key_last = twiddle_in<K>(key_last, a.tw);
      }
      uint32_t notfirst = 0;
      unsigned long long arrs = 0;               // DENSE: 4-bit arrival index of each of the first 16 keys of this thread
      uint32_t arrs_hi = 0, full15 = 0;          //        ... and of keys 16, 17 (and of the incomplete row, at field rows_full)
      const uint32_t lo_base = smem_u32(sm.lo), hi_base = smem_u32(sm.hi);
      const uint32_t outk = smem_u32(sk + aoff), outv = VB ? smem_u32(sv + voff) : 0u, org_base = smem_u32(sm.origin);
      auto mark = [&](K k, uint32_t j) {
        const uint32_t c = (uint32_t)(k >> lo_bit) & cmask;
        if (!DENSE) {
          const uint32_t b = c & 15u;
          const uint32_t old = atoms_or(lo_base + (c >> 4) * 4u, 1u << b);
          notfirst |= ((old >> b) & 1u) << j;
        } else {
          // DENSE: a 4-bit counter per cell; the atomicAdd returns the key's arrival index inside its cell (15 = the cell is full:
          // the bucket goes to the overflow list -- the add has carried into the neighbouring counter, which no longer matters)
          const uint32_t sh = (c & 7u) * 4u;
          const uint32_t a4 = (atoms_add(lo_base + (c >> 3) * 4u, 1u << sh) >> sh) & 15u;
          full15 |= a4 == 15u ? 1u : 0u;
          if (j < 16u) arrs |= (unsigned long long)a4 << (4u * j); else arrs_hi |= a4 << (4u * (j - 16u));
        }
      };
      {
        const uint32_t rf = opaque(rows_full);
#pragma unroll
        for (int j = 0; j < IPT; ++j)
          if ((uint32_t)j < rf) mark(key[j], j);
      }
      if (has_last) mark(key_last, rows_full);
      if (!DENSE) {
        for (uint32_t m = notfirst; m;) {          // the few keys whose cell was taken: second-copy bit, then the short list
          const int j = __ffs(m) - 1;
          m &= m - 1u;
          const uint32_t idx = (uint32_t)j * THREADS + tid;
          K k = sk[skew + idx];                    // (the staged keys are intact until the barrier below)
          if (a.tw_in) k = twiddle_in<K>(k, a.tw);
          const uint32_t c = (uint32_t)(k >> lo_bit) & cmask;
          const uint32_t bit = 0x10000u << (c & 15u);
          const uint32_t old = atomicOr(&sm.lo[c >> 4], bit);
          if (old & bit) {
            const uint32_t x = atomicAdd(&sm.nextra, 1u);
            if (x < (uint32_t)RANK_EXTRA) sm.extras[x] = c | (idx << 16);
          }
        }
      } else if (full15) {
        sm.nextra = RANK_EXTRA + 1u;               // any value above RANK_EXTRA: "not sorted"
      }
      __syncthreads();
      const uint32_t nx = sm.nextra;
      sorted = nx <= (uint32_t)RANK_EXTRA;
      if (sorted) {
        // ---- B: exclusive scan over the entries.  lo[] is cut into NG chunks of THREADS 16-byte vectors; thread t takes
        // vector t of every chunk (conflict-free shared-memory accesses), each chunk is scanned across the block, and the
        // NG * NWARPS = 32 (chunk, warp) totals are chained by one warp scan.
        {
          constexpr int NG = (RANK_ENT / 4) / THREADS;
          static_assert(NG * NWARPS == 32, "one lane per (chunk, warp) total");
          uint32_t pc[NG][4], sum[NG], inc[NG];
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const uint4 q = reinterpret_cast<const uint4*>(sm.lo)[g * THREADS + tid];      // (entries past nent are never set)
            if (!DENSE) { pc[g][0] = __popc(q.x); pc[g][1] = __popc(q.y); pc[g][2] = __popc(q.z); pc[g][3] = __popc(q.w); }
            else { pc[g][0] = nibble_sum(q.x); pc[g][1] = nibble_sum(q.y); pc[g][2] = nibble_sum(q.z); pc[g][3] = nibble_sum(q.w); }
          }
          uint32_t flags = 0;                    // entries of mine that hold third-or-later copies (almost always none)
          {
            constexpr int LOG_T = THREADS == 256 ? 8 : THREADS == 512 ? 9 : 10;
            static_assert((1 << LOG_T) == THREADS, "entry -> (chunk, thread, word) split");
            unsigned long long xcnt = 0;         // listed copies per entry of mine: 4-bit fields, field = chunk * 4 + word
#pragma unroll 1
            for (uint32_t i = 0; i < (DENSE ? 0u : nx); ++i) {
              const uint32_t xe = (sm.extras[i] & 0xFFFFu) >> 4;
              if (((xe >> 2) & (uint32_t)(THREADS - 1)) == tid) xcnt += 1ull << (4u * (((xe >> (2 + LOG_T)) << 2) | (xe & 3u)));
            }
            if (xcnt) {
#pragma unroll
              for (int g = 0; g < NG; ++g)
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  const uint32_t x = (uint32_t)(xcnt >> (4 * (g * 4 + i4))) & 15u;
                  pc[g][i4] += x; flags |= (x ? 1u : 0u) << (g * 4 + i4);
                }
            }
          }
#pragma unroll
          for (int g = 0; g < NG; ++g) { sum[g] = pc[g][0] + pc[g][1] + pc[g][2] + pc[g][3]; inc[g] = sum[g]; }
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
              const uint32_t t = __shfl_up_sync(0xffffffffu, inc[g], o);
              if (lane >= (unsigned)o) inc[g] += t;
            }
          }
          if (lane == 31) {
#pragma unroll
            for (int g = 0; g < NG; ++g) sm.wt[g * NWARPS + w] = inc[g];
          }
          __syncthreads();
          const uint32_t wv = sm.wt[lane];
          uint32_t wi = wv;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (unsigned)o) wi += t;
          }
          wi -= wv;
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const uint32_t p0 = __shfl_sync(0xffffffffu, wi, g * NWARPS + w) + inc[g] - sum[g];
            const uint32_t p1 = p0 + pc[g][0], p2 = p1 + pc[g][1], p3 = p2 + pc[g][2];
            uint2 out = make_uint2(p0 | (p1 << 16), p2 | (p3 << 16));
            if (flags) {
              const uint32_t f = flags >> (4 * g);
              out.x |= ((f & 1u) << 15) | ((f & 2u) << 30);
              out.y |= ((f & 4u) << 13) | ((f & 8u) << 28);
            }
            reinterpret_cast<uint2*>(sm.hi)[g * THREADS + tid] = out;
          }
        }
        __syncthreads();

        // ---- C: rank lookup and reorder into the consumed staging slot
        uint32_t pend = 0;                       // ORDER: keys of shared cells, placed after the copies have met
        uint32_t pos[ORDER ? IPT : 1], pos_last = 0;
        auto place = [&](K k, V v, uint32_t j, uint32_t& pj) {
          const uint32_t c = (uint32_t)(k >> lo_bit) & cmask;
          const uint32_t e = c >> (DENSE ? 3 : 4);
          const uint32_t l = lds_u32(lo_base + e * 4u), h = lds_u16(hi_base + e * 2u);
          uint32_t rank, arr, mult;
          if (!DENSE) {
            const uint32_t b = c & 15u;
            const uint32_t below = (0x10001u << b) - 0x10001u;        // the cells below b, first and second copies
            rank = (h & 0x7FFFu) + __popc(l & below);
            arr = (notfirst >> j) & 1u;                               // arrival order inside the cell: 0, 1, (2 + k: rare path)
            mult = ORDER ? 1u + ((l >> (16 + b)) & 1u) : 0u;
            if (h & 0x8000u) {
              // rare: the key's entry holds third-or-later copies (listed in extras[], arrival order).  Listed copies in lower cells
              // of the entry add to the rank, listed copies of the key's own cell to its multiplicity; a listed key gets its arrival index
              const uint32_t me = c | ((j * THREADS + tid) << 16);
              uint32_t same = 0;
#pragma unroll 1
              for (uint32_t i = 0; i < nx; ++i) {
                const uint32_t x = sm.extras[i], xc = x & 0xFFFFu;
                rank += ((xc ^ c) < 16u && xc < c) ? 1u : 0u;
                if (xc == c) { if (x == me) arr = 2u + same; ++same; }
              }
              mult += same;
            }
          } else {
            const uint32_t sh = (c & 7u) * 4u;
            rank = h + nibble_sum(l & ((1u << sh) - 1u));            // keys in the lower cells of the entry
            arr = j < 16u ? (uint32_t)(arrs >> (4u * j)) & 15u : (arrs_hi >> (4u * (j - 16u))) & 15u;
            mult = (l >> sh) & 15u;
          }
          if (!ORDER) {
            const uint32_t q = rank + arr;
            sts_t<K>(outk + q * (uint32_t)sizeof(K), k);
            if (VB) sts_t<V>(outv + q * (uint32_t)sizeof(V), v);
          } else if (mult == 1u) {
            sts_t<K>(outk + rank * (uint32_t)sizeof(K), k);
            if (VB) sts_t<V>(outv + rank * (uint32_t)sizeof(V), v);
          } else {
            sts_u16(org_base + (rank + arr) * 2u, j * THREADS + tid);
            pj = rank | (mult << 16);
            pend |= 1u << j;
          }
        };
        {
          const uint32_t rf = opaque(rows_full);
#pragma unroll
          for (int j = 0; j < IPT; ++j)
            if ((uint32_t)j < rf) place(key[j], val[VB ? j : 0], j, pos[ORDER ? j : 0]);
        }
        if (has_last) place(key_last, val_last, rows_full, pos_last);
        if (ORDER) {
          __syncthreads();
          auto settle = [&](K k, V v, uint32_t j, uint32_t pj) {
            const uint32_t idx = j * THREADS + tid;
            const uint32_t rank = pj & 0xFFFFu, mult = pj >> 16;
            uint32_t q = rank;
            for (uint32_t i = 0; i < mult; ++i) q += (lds_u16(org_base + (rank + i) * 2u) < idx) ? 1u : 0u;
            sts_t<K>(outk + q * (uint32_t)sizeof(K), k);
            if (VB) sts_t<V>(outv + q * (uint32_t)sizeof(V), v);
          };
#pragma unroll
          for (int j = 0; j < IPT; ++j)
            if ((uint32_t)j < rows_full && ((pend >> j) & 1u)) settle(key[j], val[VB ? j : 0], j, pos[ORDER ? j : 0]);
          if (has_last && ((pend >> rows_full) & 1u)) settle(key_last, val_last, rows_full, pos_last);
        }
      }
      __syncthreads();
      // the presence bits are no longer needed: clear them for the next bucket while the result is written out
      for (uint32_t i = tid; i < (nent + 3) / 4; i += THREADS) reinterpret_cast<uint4*>(sm.lo)[i] = make_uint4(0, 0, 0, 0);
      if (tid == 0) sm.nextra = 0;
    }
    if (!sorted && tid == 0) {
      // too many copies of some value for this variant: the dense variant takes the bucket if it can, the LSD kernel otherwise
      const bool to_dense = !DENSE && a.dense != nullptr && nb <= RANK_DENSE_BITS;
      if (to_dense) {
        const uint32_t o = atomicAdd(a.num_dense_ptr, 1u);
        if (o < a.max_items) a.dense[o] = it; else atomicOr(a.error_ptr, 2u);
      } else {
        hand_back(a, it);
      }
    }

    // ---- D: shared-memory vector v and output vector v cover the same elements, both 16-byte aligned
    if (sorted) {
      {
        K* __restrict__ gdst = keys_out + it.off - aoff;
        const uint32_t total = aoff + cnt, nv = (total + EK - 1) / EK;
        const bool two = a.tw_out != 0;
        for (uint32_t v = tid; v < nv; v += THREADS) {
          uint4 q = reinterpret_cast<const uint4*>(sk)[v];
          K* e = reinterpret_cast<K*>(&q);
          if (two) {
#pragma unroll
            for (int i = 0; i < EK; ++i) e[i] = twiddle_out<K>(e[i], a.tw);
          }
          if (v * EK >= aoff && v * EK + EK <= total) reinterpret_cast<uint4*>(gdst)[v] = q;
          else {
#pragma unroll
            for (int i = 0; i < EK; ++i)
              if (v * EK + i >= aoff && v * EK + i < total) gdst[v * EK + i] = e[i];
          }
        }
      }
      if constexpr (VB != 0) {
        V* __restrict__ gdst = vals_out + it.off - voff;
        const uint32_t total = voff + cnt, nv = (total + EV - 1) / EV;
        for (uint32_t v = tid; v < nv; v += THREADS) {
          const uint4 q = reinterpret_cast<const uint4*>(sv)[v];
          const V* e = reinterpret_cast<const V*>(&q);
          if (v * EV >= voff && v * EV + EV <= total) reinterpret_cast<uint4*>(gdst)[v] = q;
          else {
#pragma unroll
            for (int i = 0; i < EV; ++i)
              if (v * EV + i >= voff && v * EV + i < total) gdst[v * EV + i] = e[i];
          }
        }
      }
    }
    __syncthreads();
    if (tid == PRODUCER) it_a = it_b;
  }
}

}  // namespace b200
