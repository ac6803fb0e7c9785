// scatter.cuh -- the bucket scatter: one read + one write sweep that splits every segment by one digit.
//
// One kernel template serves
//   MODE_SEG   : every level of the MSB hybrid sort (many segments; tiles come from a device-built TileDesc list; a
//                tile reserves its output chunk per digit with one global atomicAdd, so tiles are independent and the
//                order inside a sub-bucket is arbitrary -- exactly the freedom the reference takes,
//                msb/src/sort/cuda_radix_sort.h:408-417).  Replaces rdxsrt_partition_keys (cuda_radix_sort.h:363-479).
//   MODE_LSB   : every pass of the stable LSB sort (one segment, onesweep style: digit starts from the up-front
//                histogram, tile prefix by decoupled look-back, stable in-tile ranking).  Replaces
//                DeviceRadixSortDownsweepKernel (lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:164-196).
//   MODE_RANGE : the multi-GPU send partition: MODE_SEG addressing (one segment, per-tile counts), stable, digit =
//                destination rank by key range; every destination may live in its own buffer (peer GPU memory).
//
// The kernel is bound by instruction issue and shared-memory traffic long before HBM (DESIGN.md "Kernels"), so the
// per-key path is kept minimal: full tiles take a path without any bounds checks, the order-preserving transform
// is applied in its own uniform-branch loop only on the first / last sweep, destinations are per-digit 64-bit
// pointers in shared memory (one LDS.64 + one IMAD.WIDE per key), and the digit is a 2-instruction shift+mask.
//
// Structure (persistent CTAs):
//   * the NEXT tile's keys (and values) are staged into shared memory by TMA bulk copies (cp.async.bulk + mbarrier)
//     issued by one producer thread while the current tile is being ranked;
//   * keys are ranked from registers: MSB -> one shared-memory atomicAdd-with-return per key; LSB -> lanes holding the
//     same digit find each other through an atomicOr of their lane bit into a per-warp match mask and rank against
//     warp-private running counters (tools/ubench_rank.cu measured the candidates on B200);
//   * digit owners (threads 0..255) reserve / look back and publish one destination pointer per digit;
//   * keys (and values) are reordered through the consumed staging buffer so every digit's run leaves as consecutive
//     addresses (coalesced stores; adjacent tiles complete each other's partial sectors in the 126 MB L2).
#pragma once
#include "async.cuh"
#include "common.cuh"
#include "hist.cuh"

namespace b200 {

enum { MODE_SEG = 0, MODE_LSB = 1, MODE_RANGE = 2 };

struct ScatterArgs {
  const void* keys_in; void* keys_out;
  const void* vals_in; void* vals_out;
  const TileDesc* descs;          // MODE_SEG: tile -> (offset, count, segment, tile in segment)
  const uint32_t* num_tiles_ptr;  // MODE_SEG: device-side tile count
  uint32_t num_tiles;             // MODE_LSB
  uint64_t base, n;               // MODE_LSB: the launch covers keys [base, base+n)
  const uint64_t* bins;           // MODE_SEG: [segment][256] absolute output index of the start of each sub-bucket;
                                  // otherwise [256] absolute output index of the start of each digit
  const uint32_t* tile_off;       // MODE_SEG: [tile][256]  keys of the same (segment, digit) in earlier tiles of the tile's group
  const uint32_t* carry;          // MODE_SEG: [group][256] ... and in earlier groups (tile_hist_kernel / group_carry_kernel)
  const uint16_t* tile_cnt;       // scatter_fast_kernel: [tile][256] keys of every digit in the tile
  uint64_t* bins_next;            // MODE_LSB: the last tile writes bins + portion counts here (or nullptr)
  uint32_t* status;               // MODE_LSB: [tile][256] look-back words, zeroed before the launch
  uint32_t* ticket;               // MODE_LSB: zeroed before the launch
  int shift; uint32_t mask;
  int tw_in, tw_out;
  Twiddle tw;
  const uint32_t* splitters; int num_parts;     // MODE_RANGE: digit = #{ j < num_parts-1 : splitters[j] <= (key >> shift) }
  const uint64_t* dst_keys; const uint64_t* dst_vals;   // MODE_RANGE: [num_parts] base ADDRESS of every destination's key / value buffer
                                                        // (nullptr: keys_out / vals_out for all); bins[d] = first index inside that buffer
  const uint8_t* digit_dest;      // scatter_stable_fast_kernel<PEER> (multi-GPU exchange): [256] which of the dst_keys / dst_vals buffers
                                  // (the peer GPUs' receive buffers) an exchange BUCKET goes to; bucket = digit >> xshift (several digits
                                  // share a bucket so that a tile's run for a bucket is long enough for NVLink: >= 128-byte pieces reach
                                  // 700 GB/s, 64-byte pieces 420, profiles/r02_ubench_peer.jsonl); bins[b] = the bucket's first index there
  int xshift;
};

constexpr int MAX_PARTS = 16;

struct TileGeom {     // per staging slot, written by the producer thread
  uint64_t off;       // first key index
  uint32_t cnt, seg, tile_in_seg, tile, skew, vskew;
};

template <typename K, int VB, int THREADS, int IPT, int MODE, bool ORD>
struct ScatterSmem {
  static constexpr int TILE = THREADS * IPT;
  static constexpr int WARPS = THREADS / 32;
  static constexpr bool ORDERED = MODE == MODE_LSB || ORD;
  using V = typename ValType<VB>::type;
  static constexpr int SLACK = 16 / sizeof(K), VSLACK = 16 / sizeof(V);
  alignas(16) K stage[2][TILE + SLACK];
  alignas(16) V vstage[VB ? 2 : 1][VB ? TILE + VSLACK : 1];
  K* kptr[RADIX];                                   // per digit: keys_out + (global start - start inside the tile)
  V* vptr[VB ? RADIX : 1];
  alignas(16) uint32_t match[ORDERED ? 2 : 1][ORDERED ? WARPS * RADIX : 1];   // per-warp match masks, two alternating sets
  alignas(16) uint16_t wcnt[ORDERED ? WARPS * RADIX : 2];                     // per-warp counters, later per-warp start positions
  uint32_t cnt[RADIX];                              // MSB: tile histogram (atomic ranking counters)
  uint32_t bin_start[RADIX];                        // tile-local exclusive start of each digit
  uint32_t scratch[8];
  alignas(8) uint64_t bar[2];
  TileGeom geom[2];
  RangeLut range;
  uint32_t skewed;                                  // MSB: the previous tile had a dominant digit -> aggregate per warp
};

template <typename K, int MODE>
__device__ __forceinline__ uint32_t scatter_digit(K k, int shift, uint32_t mask, const RangeLut& rl, int cshift) {
  if (MODE != MODE_RANGE) return digit_of<K>(k, shift, mask);
  return range_part(rl, (uint32_t)(k >> shift), cshift);
}

// 32-bit / 64-bit forms of the order-preserving transform with the masks already narrowed to K (3 instructions).
template <typename K>
__device__ __forceinline__ K tw_apply_in(K k, K sign, K fl, K flip) {
  using S = typename std::make_signed<K>::type;
  return (K)(k ^ (((K)((S)k >> (sizeof(K) * 8 - 1)) & fl) | sign) ^ flip);
}
template <typename K>
__device__ __forceinline__ K tw_apply_out(K k, K sign, K fl, K flip) {
  using S = typename std::make_signed<K>::type;
  k = (K)(k ^ flip);
  return (K)(k ^ (((K)(~(K)((S)k >> (sizeof(K) * 8 - 1))) & fl) | sign));
}

// Store through a per-digit destination pointer kept in shared memory (known to be global memory: plain STG, no
// generic-address resolution).
template <typename T>
__device__ __forceinline__ void st_global(T* base, uint32_t idx, T v) {
  T* p = base + idx;
  if (sizeof(T) == 4) asm volatile("st.global.b32 [%0], %1;" ::"l"(p), "r"((uint32_t)v) : "memory");
  else asm volatile("st.global.b64 [%0], %1;" ::"l"(p), "l"((uint64_t)v) : "memory");
}

#ifndef LB_BATCH
#define LB_BATCH 8
#endif
#ifndef LB_POS
#define LB_POS 0
#endif
// Decoupled look-back for one digit (called by the digit's owner thread): sums the aggregates of the predecessor tiles
// down to the nearest one that already knows its inclusive prefix, LB_BATCH predecessors per round trip (the loads of
// a batch are independent, so a walk of depth D costs ~D/LB_BATCH L2 latencies).  A stale word is still valid (an
// aggregate is only ever upgraded to a prefix), so only not-yet-published words are re-read.  Publishes this tile's
// inclusive prefix and returns its exclusive prefix.
__device__ __forceinline__ uint32_t lookback(const uint32_t* sbase, uint32_t tile, uint32_t tile_in_seg, uint32_t* stw, uint32_t my_total) {
  uint32_t excl_g = 0;
  int64_t t = (int64_t)tile - 1;
  const int64_t t_first = (int64_t)tile - (int64_t)tile_in_seg;
  bool done = false;
  while (!done) {
    uint32_t s[LB_BATCH];
#pragma unroll
    for (int j = 0; j < LB_BATCH; ++j) s[j] = (t - j >= t_first) ? ld_status(sbase + (uint64_t)(t - j) * RADIX) : ST_PREFIX;
#pragma unroll
    for (int j = 0; j < LB_BATCH; ++j) {
      if (!done) {
        uint32_t v = s[j];
        while ((v >> 30) == 0) { __nanosleep(20); v = ld_status(sbase + (uint64_t)(t - j) * RADIX); }
        excl_g += v & ST_VALUE_MASK;
        done = (v & ST_PREFIX) != 0;
      }
    }
    t -= LB_BATCH;
  }
  st_status(stw, ST_PREFIX | (excl_g + my_total));
  return excl_g;
}

// ---------------------------------------------------------------------------------------------------------------
// One tile.  FULL = the tile holds exactly TILE keys (no bounds checks anywhere).
// ---------------------------------------------------------------------------------------------------------------
template <typename K, int VB, int THREADS, int IPT, int MODE, bool ORD, bool FULL>
__device__ __forceinline__ void scatter_tile(const ScatterArgs& a, ScatterSmem<K, VB, THREADS, IPT, MODE, ORD>& sm, const TileGeom& g, int slot,
                                             uint32_t num_tiles, uint32_t it) {
  using SM = ScatterSmem<K, VB, THREADS, IPT, MODE, ORD>;
  using V = typename SM::V;
  constexpr int TILE = SM::TILE, WARPS = SM::WARPS;
  constexpr bool ORDERED = SM::ORDERED;
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const uint32_t cnt = FULL ? (uint32_t)TILE : g.cnt;
  const int shift = a.shift; const uint32_t mask = a.mask;
  const int cshift = MODE == MODE_RANGE ? sm.range.cshift : 0;
  K* __restrict__ st = &sm.stage[slot][0];
  V* __restrict__ vst = &sm.vstage[VB ? slot : 0][0];

  // ---- keys of this tile: shared memory (TMA-staged) -> registers
  mbar_wait(&sm.bar[slot], (it >> 1) & 1u);
  K key[IPT];
  const uint32_t ibase = ORDERED ? w * (32u * IPT) + lane : tid;     // index of item j: ibase + j * istep
  constexpr uint32_t istep = ORDERED ? 32u : (uint32_t)THREADS;
  {
    const K* __restrict__ src = st + g.skew + ibase;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      if (FULL) key[j] = src[j * istep];
      else key[j] = (ibase + j * istep < cnt) ? src[j * istep] : (K)~(K)0;
    }
  }
  if (a.tw_in) {
    const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      if (FULL) key[j] = tw_apply_in<K>(key[j], sg, fl, fp);
      else key[j] = (ibase + j * istep < cnt) ? tw_apply_in<K>(key[j], sg, fl, fp) : (K)~(K)0;
    }
  }

  // ---- rank inside the tile
  uint32_t pos[IPT];
  uint32_t my_total = 0;
  if (!ORDERED) {
    if (tid < RADIX) sm.cnt[tid] = 0;
    __syncthreads();
    if (!sm.skewed) {
#pragma unroll
      for (int j = 0; j < IPT; ++j)
        if (FULL || ibase + j * istep < cnt) pos[j] = atomicAdd(&sm.cnt[digit_of<K>(key[j], shift, mask)], 1u);
    } else {
      // dominant digit: same-address shared-memory atomics with return serialise (~11x slower on constant input,
      // profiles/ubench_rank_r01.jsonl) -> a warp whose 32 digits agree adds once for everybody
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        const bool v = FULL || ibase + j * istep < cnt;
        const unsigned d = digit_of<K>(key[j], shift, mask);
        const unsigned d0 = __shfl_sync(0xffffffffu, d, 0);
        if (__all_sync(0xffffffffu, v && d == d0)) {
          unsigned b = 0;
          if (lane == 0) b = atomicAdd(&sm.cnt[d0], 32u);
          pos[j] = __shfl_sync(0xffffffffu, b, 0) + lane;
        } else if (v) {
          pos[j] = atomicAdd(&sm.cnt[d], 1u);
        }
      }
    }
    __syncthreads();
    if (tid < RADIX) my_total = sm.cnt[tid];
  } else {
    // stable ranking.  Row j of a warp = its 32 keys j*32 .. j*32+31 of the warp's contiguous share; rows are ranked in
    // order, lanes in order inside a row.  Lanes with equal digits meet in match[j&1][w][d] (atomicOr of the lane bit);
    // the lowest such lane adds the row's count to the warp's running counter and hands the old value to its peers.
    // (the match masks are all zero here: zeroed once at kernel start, and every row's leader clears the word it used)
    uint4* zc = reinterpret_cast<uint4*>(sm.wcnt);
    for (int i = tid; i < WARPS * RADIX / 8; i += THREADS) zc[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    uint16_t* wc = sm.wcnt + w * RADIX;
    const unsigned lt = (1u << lane) - 1u, lbit = 1u << lane;
    if (MODE == MODE_RANGE) {
      // few destinations: same-address atomics would serialise, so the lanes of a row find their peers with one ballot per
      // destination instead (uniform loop over num_parts <= 16)
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        const unsigned d = scatter_digit<K, MODE>(key[j], shift, mask, sm.range, cshift);
        unsigned peers = 0;
        for (int p = 0; p < a.num_parts; ++p) {
          const unsigned m = __ballot_sync(0xffffffffu, d == (unsigned)p);
          if (d == (unsigned)p) peers = m;
        }
        const unsigned below = __popc(peers & lt);
        unsigned b = 0;
        if (below == 0) { b = wc[d]; wc[d] = (uint16_t)(b + __popc(peers)); }
        b = __shfl_sync(0xffffffffu, b, __ffs(peers) - 1);
        pos[j] = b + below;
        __syncwarp();
      }
    } else {
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      uint32_t* wm = sm.match[j & 1] + w * RADIX;
      const unsigned d = scatter_digit<K, MODE>(key[j], shift, mask, sm.range, cshift);   // padding keys (all ones) rank last
      atomicOr(&wm[d], lbit);
      __syncwarp();
      const unsigned peers = wm[d];
      __syncwarp();
      const unsigned below = __popc(peers & lt);
      unsigned b = 0;
      if (below == 0) { b = wc[d]; wc[d] = (uint16_t)(b + __popc(peers)); wm[d] = 0; }
      b = __shfl_sync(0xffffffffu, b, __ffs(peers) - 1);
      pos[j] = b + below;
    }
    }
    __syncthreads();
    if (tid < RADIX) {
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) my_total += sm.wcnt[ww * RADIX + tid];
      if (!FULL && tid == scatter_digit<K, MODE>((K)~(K)0, shift, mask, sm.range, cshift)) my_total -= (uint32_t)TILE - cnt;   // padding
    }
  }

  // ---- digit owners: reserve (MSB) / publish the aggregate (LSB) as early as possible, then the tile-local scan
  uint64_t gstart = 0;
  uint32_t excl_g = 0;
  uint32_t* stw = nullptr;
  const bool first = g.tile_in_seg == 0;
  if (tid < RADIX) {
    if (MODE != MODE_LSB) {
      // destination = start of the (segment, digit) sub-bucket + keys of it in earlier tiles: three independent loads whose
      // latency hides behind the scan and the shared-memory reorder below
      const uint32_t grp = g.tile / HIST_GROUP;
      const bool continues = g.tile - g.tile_in_seg < grp * HIST_GROUP;       // the segment started in an earlier group
      gstart = a.bins[(uint64_t)g.seg * RADIX + tid] + a.tile_off[(uint64_t)g.tile * RADIX + tid];
      if (continues) gstart += a.carry[(uint64_t)grp * RADIX + tid];
    } else {
      stw = a.status + (uint64_t)g.tile * RADIX + tid;
      st_status(stw, (first ? ST_PREFIX : ST_AGG) | my_total);
      if (LB_POS == 1 && !first) excl_g = lookback(a.status + tid, g.tile, g.tile_in_seg, stw, my_total);
    }
  }
  uint32_t inc = my_total;
  if (tid < RADIX) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) sm.scratch[w] = inc;
  }
  __syncthreads();
  uint32_t my_excl = 0;
  if (tid < RADIX) {
    uint32_t woff = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) woff += ((unsigned)j < w) ? sm.scratch[j] : 0u;
    my_excl = woff + inc - my_total;
    if (!ORDERED) {
      sm.bin_start[tid] = my_excl;
    } else {
      uint32_t run = my_excl;
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) {
        const uint32_t c = sm.wcnt[ww * RADIX + tid];
        sm.wcnt[ww * RADIX + tid] = (uint16_t)run;
        run += c;
      }
    }
  }
  __syncthreads();

  // ---- reorder through shared memory (every thread has its keys in registers: the staging buffer is reused)
  if (!ORDERED) {
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if (FULL || ibase + j * istep < cnt) { pos[j] += sm.bin_start[digit_of<K>(key[j], shift, mask)]; st[pos[j]] = key[j]; }
  } else {
    const uint16_t* wc = sm.wcnt + w * RADIX;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      pos[j] += wc[scatter_digit<K, MODE>(key[j], shift, mask, sm.range, cshift)];
      if (FULL || pos[j] < cnt) st[pos[j]] = key[j];        // padding keys rank after every real key: pos >= cnt
    }
  }
  V val[VB ? IPT : 1];
  if (VB) {
    const V* __restrict__ vsrc = vst + g.vskew + ibase;
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if (FULL || ibase + j * istep < cnt) val[j] = vsrc[j * istep];
  }

  // ---- digit owners: global start of this tile's run of every digit -> destination pointers
  if (tid < RADIX) {
    if (MODE == MODE_LSB) {
      if (LB_POS == 0) excl_g = first ? 0u : lookback(a.status + tid, g.tile, g.tile_in_seg, stw, my_total);
      gstart = a.bins[tid] + excl_g;
      if (a.bins_next != nullptr && g.tile == num_tiles - 1) a.bins_next[tid] = gstart + my_total;
    }
    K* kbase = reinterpret_cast<K*>(a.keys_out);
    V* vbase = reinterpret_cast<V*>(a.vals_out);
    if (MODE == MODE_RANGE && a.dst_keys != nullptr && (int)tid < a.num_parts) {       // per-destination buffers (peer memory)
      kbase = reinterpret_cast<K*>(a.dst_keys[tid]);
      if (VB) vbase = reinterpret_cast<V*>(a.dst_vals[tid]);
    }
    sm.kptr[tid] = kbase + (gstart - my_excl);
    if (VB) sm.vptr[tid] = vbase + (gstart - my_excl);
  }
  __syncthreads();        // keys are in place, every thread has read its values, kptr/vptr are published
  if (VB) {
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if (FULL || (ORDERED ? pos[j] < cnt : ibase + j * istep < cnt)) vst[pos[j]] = val[j];
    __syncthreads();
  }

  // ---- coalesced write-out: consecutive positions of one digit are consecutive output addresses
  if (a.tw_out) {
    const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const uint32_t p = j * THREADS + tid;
      if (FULL || p < cnt) {
        const K k = st[p];
        const uint32_t d = scatter_digit<K, MODE>(k, shift, mask, sm.range, cshift);
        st_global<K>(sm.kptr[d], p, tw_apply_out<K>(k, sg, fl, fp));
        if (VB) st_global<V>(sm.vptr[d], p, vst[p]);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const uint32_t p = j * THREADS + tid;
      if (FULL || p < cnt) {
        const K k = st[p];
        const uint32_t d = scatter_digit<K, MODE>(k, shift, mask, sm.range, cshift);
        st_global<K>(sm.kptr[d], p, k);
        if (VB) st_global<V>(sm.vptr[d], p, vst[p]);
      }
    }
  }
  if (!ORDERED) {
    // the next tile aggregates per warp if this one had a dominant digit (tiles of a CTA are neighbours in key space)
    const int sk = __syncthreads_or(my_total > (uint32_t)TILE / 4);
    if (tid == 0) sm.skewed = sk ? 1u : 0u;
  } else {
    __syncthreads();
  }
}

template <typename K, int VB, int THREADS, int IPT, int OCC, int MODE, bool ORD>
__global__ void __launch_bounds__(THREADS, OCC) scatter_kernel(const __grid_constant__ ScatterArgs a) {
  using SM = ScatterSmem<K, VB, THREADS, IPT, MODE, ORD>;
  using V = typename SM::V;
  constexpr int TILE = SM::TILE;
  constexpr bool ORDERED = SM::ORDERED;
  constexpr unsigned PRODUCER = THREADS - 1;
  static_assert(THREADS >= RADIX, "one digit owner per digit");
  static_assert(TILE < 65536, "per-warp start positions are 16-bit");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x;
  const K* __restrict__ keys_in = reinterpret_cast<const K*>(a.keys_in);
  const V* __restrict__ vals_in = reinterpret_cast<const V*>(a.vals_in);
  const uint32_t num_tiles = MODE != MODE_LSB ? *a.num_tiles_ptr : a.num_tiles;

  // producer: describe tile `t`, arm the slot's barrier and launch the bulk copies of its keys (and values)
  auto stage_tile = [&](int slot, uint32_t t, const TileDesc& td) {
    TileGeom g;
    g.tile = t; g.skew = 0; g.vskew = 0;
    if (t < num_tiles) {
      if (MODE != MODE_LSB) { g.off = td.off; g.cnt = td.cnt; g.seg = td.seg; g.tile_in_seg = td.tile_in_seg; }
      else {
        const uint64_t rel = (uint64_t)t * TILE;
        g.off = a.base + rel; g.cnt = (uint32_t)(a.n - rel < (uint64_t)TILE ? a.n - rel : (uint64_t)TILE);
        g.seg = 0; g.tile_in_seg = t;
      }
      const BulkWindow<K> bw(keys_in, g.off, g.cnt);
      g.skew = bw.skew;
      uint32_t bytes = bw.bytes;
      fence_proxy_async();
      if (VB) {
        const BulkWindow<V> vw(vals_in, g.off, g.cnt);
        g.vskew = vw.skew;
        bytes += vw.bytes;
        mbar_expect_tx(&sm.bar[slot], bytes);
        bulk_g2s(&sm.vstage[VB ? slot : 0][0], vw.src, vw.bytes, &sm.bar[slot]);
      } else {
        mbar_expect_tx(&sm.bar[slot], bytes);
      }
      bulk_g2s(&sm.stage[slot][0], bw.src, bw.bytes, &sm.bar[slot]);
    } else {
      g.off = 0; g.cnt = 0; g.seg = 0; g.tile_in_seg = 0;
    }
    sm.geom[slot] = g;
  };

  // Tile sequence.  Ordered modes: dynamic tickets; MSB: tiles are independent and dealt round-robin.  The producer
  // knows its next tile one full iteration ahead (ticket / descriptor fetched during the previous tile), so the TMA
  // prefetch into the free slot goes out at the very top of an iteration.  Holding tickets ahead cannot deadlock the
  // look-back: a CTA processes its tiles in increasing order, hence the lowest unfinished tile of the launch is always
  // the one its CTA is working on, it only waits for lower (finished) tiles, and all CTAs are co-resident.
  uint32_t tk_a = 0, tk_b = 0;      // producer only: tile of the next / the following iteration
  TileDesc td_a{}, td_b{};
  if (tid == PRODUCER) {
    mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
    const uint32_t t0 = MODE == MODE_LSB ? atomicAdd(a.ticket, 1u) : blockIdx.x;
    tk_a = MODE == MODE_LSB ? atomicAdd(a.ticket, 1u) : t0 + gridDim.x;
    TileDesc td{};
    if (MODE != MODE_LSB && t0 < num_tiles) td = a.descs[t0];
    if (MODE != MODE_LSB && tk_a < num_tiles) td_a = a.descs[tk_a];
    stage_tile(0, t0, td);
  }
  if (tid == 0) sm.skewed = 0;
  if (ORDERED) {
    uint4* z = reinterpret_cast<uint4*>(sm.match);
    for (int i = tid; i < 2 * SM::WARPS * RADIX / 4; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (MODE == MODE_RANGE) range_lut_build(sm.range, a.splitters, a.num_parts, (int)sizeof(K) * 8 - a.shift);
  __syncthreads();

  for (uint32_t it = 0;; ++it) {
    const int slot = (int)(it & 1u);
    const TileGeom g = sm.geom[slot];
    if (g.tile >= num_tiles) break;
    if (tid == PRODUCER) {
      stage_tile(slot ^ 1, tk_a, td_a);          // the other slot is free: its tile finished last iteration
      tk_b = MODE == MODE_LSB ? atomicAdd(a.ticket, 1u) : tk_a + gridDim.x;
      if (MODE != MODE_LSB && tk_b < num_tiles) td_b = a.descs[tk_b];
    }
    if (g.cnt == (uint32_t)TILE) scatter_tile<K, VB, THREADS, IPT, MODE, ORD, true>(a, sm, g, slot, num_tiles, it);
    else scatter_tile<K, VB, THREADS, IPT, MODE, ORD, false>(a, sm, g, slot, num_tiles, it);
    if (tid == PRODUCER) { tk_a = tk_b; td_a = td_b; }
  }
}

// ===============================================================================================================
// scatter_fast_kernel -- the unstable segment scatter (MODE_SEG, atomic ranking) with everything a tile needs from global
// memory fetched ONE TILE AHEAD by the digit-owner threads: the tile's digit counts (tile_hist_kernel wrote them), their
// exclusive scan and the global destinations.  The ranking counters start at each digit's first slot, so a key's atomicAdd
// returns its final position in the reorder buffer directly: no per-key bin_start lookup, no post-rank scan, and two block
// barriers per tile instead of six (the shared-memory pipe and barrier stalls bound the general kernel, profiles/r01_full_cfg2.txt).
// ===============================================================================================================
template <typename K, int VB, int THREADS, int IPT>
struct FastSmem {
  static constexpr int TILE = THREADS * IPT;
  using V = typename ValType<VB>::type;
  static constexpr int SLACK = 16 / sizeof(K), VSLACK = 16 / sizeof(V);
  alignas(16) K stage[2][TILE + SLACK];
  alignas(16) V vstage[VB ? 2 : 1][VB ? TILE + VSLACK : 1];
  uint32_t goff[2][RADIX];          // per digit: (global start - start inside the tile) mod 2^32; output index = goff[d] + position
  uint32_t cnt[2][RADIX];           // ranking counters, preset to the first slot of every digit
  uint32_t scratch[2][8];
  alignas(8) uint64_t bar[2];
  TileGeom geom[2];
  uint32_t skewed[2];
  uint32_t hot[2];                  // a digit that holds more than an eighth of the tile (the largest such digit number), when skewed
};

template <typename K, int VB, int THREADS, int IPT, int OCC>
__global__ void __launch_bounds__(THREADS, OCC) scatter_fast_kernel(const __grid_constant__ ScatterArgs a) {
  pdl_wait();
  using SM = FastSmem<K, VB, THREADS, IPT>;
  using V = typename SM::V;
  constexpr int TILE = SM::TILE;
  constexpr unsigned PRODUCER = THREADS - 1;
  static_assert(THREADS >= 2 * RADIX, "the digit owners (warps 0-7) must not include the producer's warp");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const K* __restrict__ keys_in = reinterpret_cast<const K*>(a.keys_in);
  const V* __restrict__ vals_in = reinterpret_cast<const V*>(a.vals_in);
  const uint32_t num_tiles = *a.num_tiles_ptr;
  const int shift = a.shift; const uint32_t mask = a.mask;

  auto stage_tile = [&](int slot, uint32_t t, const TileDesc& td) {
    TileGeom g;
    g.tile = t; g.skew = 0; g.vskew = 0;
    if (t < num_tiles) {
      g.off = td.off; g.cnt = td.cnt; g.seg = td.seg; g.tile_in_seg = td.tile_in_seg;
      const BulkWindow<K> bw(keys_in, g.off, g.cnt);
      g.skew = bw.skew;
      uint32_t bytes = bw.bytes;
      fence_proxy_async();
      if (VB) {
        const BulkWindow<V> vw(vals_in, g.off, g.cnt);
        g.vskew = vw.skew;
        bytes += vw.bytes;
        mbar_expect_tx(&sm.bar[slot], bytes);
        bulk_g2s(&sm.vstage[VB ? slot : 0][0], vw.src, vw.bytes, &sm.bar[slot]);
      } else {
        mbar_expect_tx(&sm.bar[slot], bytes);
      }
      bulk_g2s(&sm.stage[slot][0], bw.src, bw.bytes, &sm.bar[slot]);
    } else {
      g.off = 0; g.cnt = 0; g.seg = 0; g.tile_in_seg = 0;
    }
    sm.geom[slot] = g;
  };
  // digit owners (threads 0..255): counts, scan and destinations of the tile described by geom[slot] -> cnt / kptr / vptr[slot].
  // Two steps: prepare_load issues the global loads (their latency then hides behind the owner's share of the write-out),
  // prepare_finish consumes them.
  struct Prep { uint32_t c; uint64_t gstart; bool live; };
  auto prepare_load = [&](int slot) -> Prep {
    Prep p; p.c = 0; p.gstart = 0;
    const TileGeom g = sm.geom[slot];
    p.live = g.tile < num_tiles;                           // uniform over the 256 digit owners
    if (!p.live) return p;
    p.c = a.tile_cnt[(uint64_t)g.tile * RADIX + tid];
    const uint32_t grp = g.tile / HIST_GROUP;
    p.gstart = a.bins[(uint64_t)g.seg * RADIX + tid] + a.tile_off[(uint64_t)g.tile * RADIX + tid];
    if (g.tile - g.tile_in_seg < grp * HIST_GROUP) p.gstart += a.carry[(uint64_t)grp * RADIX + tid];     // the segment started in an earlier group
    return p;
  };
  auto prepare_finish = [&](int slot, const Prep& p) {
    if (!p.live) return;
    const uint32_t c = p.c; const uint64_t gstart = p.gstart;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) sm.scratch[slot][w] = inc;
    if (tid == 0) { sm.skewed[slot] = 0; sm.hot[slot] = 0; }
    asm volatile("bar.sync 1, 256;" ::: "memory");          // the eight digit-owner warps only
    uint32_t woff = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) woff += ((unsigned)j < w) ? sm.scratch[slot][j] : 0u;
    const uint32_t excl = woff + inc - c;
    sm.cnt[slot][tid] = VB == 0 ? smem_u32(&sm.stage[slot][0]) + excl * (uint32_t)sizeof(K) : excl;      // keys-only: shared ADDRESS of the digit's first slot
    sm.goff[slot][tid] = (uint32_t)gstart - excl;          // n < 2^32: indices wrap correctly in 32 bits
    if (c > (uint32_t)TILE / 8) { sm.skewed[slot] = 1; atomicMax(&sm.hot[slot], tid); }      // a dominant digit: its lanes share one atomic per warp and row
  };

  uint32_t tk_a = 0, tk_b = 0;
  TileDesc td_a{}, td_b{};
  if (tid == PRODUCER) {
    mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
    const uint32_t t0 = blockIdx.x;
    tk_a = t0 + gridDim.x;
    TileDesc td{};
    if (t0 < num_tiles) td = a.descs[t0];
    if (tk_a < num_tiles) td_a = a.descs[tk_a];
    stage_tile(0, t0, td);
  }
  __syncthreads();
  if (tid < RADIX) prepare_finish(0, prepare_load(0));
  __syncthreads();

  for (uint32_t it = 0;; ++it) {
    const int slot = (int)(it & 1u);
    const TileGeom g = sm.geom[slot];
    if (g.tile >= num_tiles) break;
    const uint32_t cnt = g.cnt;
    const bool full = cnt == (uint32_t)TILE;
    if (tid == PRODUCER) {
      stage_tile(slot ^ 1, tk_a, td_a);          // the other slot is free: its tile finished last iteration
      tk_b = tk_a + gridDim.x;
      if (tk_b < num_tiles) td_b = a.descs[tk_b];
    }
    K* __restrict__ st = &sm.stage[slot][0];
    V* __restrict__ vst = &sm.vstage[VB ? slot : 0][0];
    uint32_t* __restrict__ ctr = sm.cnt[slot];

    if constexpr (VB == 0) {
      // ---- keys: shared memory (TMA-staged) -> registers.  All per-key shared-memory traffic below goes through explicit
      // 32-bit shared addresses (async.cuh): the ranking counters hold the ADDRESS of every digit's next slot, so a key's
      // atomicAdd returns where it goes and the reorder store needs no address arithmetic at all.
      mbar_wait(&sm.bar[slot], (it >> 1) & 1u);
      K key[IPT]; uint32_t pos[IPT];
      const uint32_t st_base = smem_u32(st), vst_base = smem_u32(vst), ctr_base = smem_u32(ctr), go_base = smem_u32(sm.goff[slot]);
      {
        const uint32_t src = st_base + (g.skew + tid) * (uint32_t)sizeof(K);
        if (full) {
  #pragma unroll
          for (int j = 0; j < IPT; ++j) key[j] = lds_t<K>(src + j * THREADS * (uint32_t)sizeof(K));
        } else {
  #pragma unroll
          for (int j = 0; j < IPT; ++j) key[j] = (tid + j * THREADS < cnt) ? lds_t<K>(src + j * THREADS * (uint32_t)sizeof(K)) : (K)~(K)0;
        }
      }
      if (a.tw_in) {
        const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
  #pragma unroll
        for (int j = 0; j < IPT; ++j) key[j] = tw_apply_in<K>(key[j], sg, fl, fp);
      }
      V val[VB ? IPT : 1];
      if (VB) {
        const uint32_t vsrc = vst_base + (g.vskew + tid) * (uint32_t)sizeof(V);
  #pragma unroll
        for (int j = 0; j < IPT; ++j)
          if (full || tid + j * THREADS < cnt) val[j] = lds_t<V>(vsrc + j * THREADS * (uint32_t)sizeof(V));
      }
      // ---- rank: the atomicAdd returns the shared address of the key's final slot in the reorder buffer
      if (!sm.skewed[slot]) {
        if (full) {
  #pragma unroll
          for (int j = 0; j < IPT; ++j) pos[j] = atoms_add(ctr_base + digit_of<K>(key[j], shift, mask) * 4u, (uint32_t)sizeof(K));
        } else {
  #pragma unroll
          for (int j = 0; j < IPT; ++j)
            if (tid + j * THREADS < cnt) pos[j] = atoms_add(ctr_base + digit_of<K>(key[j], shift, mask) * 4u, (uint32_t)sizeof(K));
        }
      } else {
        // a dominant digit: the lanes holding it share ONE atomic per warp and row (same-address atomics serialise)
        unsigned hot_d = sm.hot[slot];
        const unsigned lt = (1u << lane) - 1u;
  #pragma unroll
        for (int j = 0; j < IPT; ++j) {
          const bool v = tid + j * THREADS < cnt;
          const unsigned d = digit_of<K>(key[j], shift, mask);
          const unsigned hot = __ballot_sync(0xffffffffu, v && d == hot_d);
          if (v && d == hot_d) {
            unsigned b = 0;
            if ((hot & lt) == 0u) b = atoms_add(ctr_base + d * 4u, (uint32_t)__popc(hot) * (uint32_t)sizeof(K));
            pos[j] = __shfl_sync(hot, b, __ffs(hot) - 1) + (uint32_t)__popc(hot & lt) * (uint32_t)sizeof(K);
          } else if (v) {
            pos[j] = atoms_add(ctr_base + d * 4u, (uint32_t)sizeof(K));
          }
        }
      }
      __syncthreads();          // every thread has its keys (and values) in registers: the staging buffers become the reorder buffers
      {
        // value slot of a key: same element index as its key slot
        auto vaddr = [&](uint32_t kaddr) { return sizeof(V) == sizeof(K) ? kaddr + (vst_base - st_base) : vst_base + ((kaddr - st_base) / (uint32_t)sizeof(K)) * (uint32_t)sizeof(V); };
        if (full) {
  #pragma unroll
          for (int j = 0; j < IPT; ++j) { sts_t<K>(pos[j], key[j]); if (VB) sts_t<V>(vaddr(pos[j]), val[j]); }
        } else {
  #pragma unroll
          for (int j = 0; j < IPT; ++j)
            if (tid + j * THREADS < cnt) { sts_t<K>(pos[j], key[j]); if (VB) sts_t<V>(vaddr(pos[j]), val[j]); }
        }
      }
      __syncthreads();          // reorder complete; geom[slot ^ 1] (written by the producer above) is visible
      // ---- digit owners fetch the next tile's counts and destinations; the loads complete behind the write-out
      Prep prep{}; if (tid < RADIX) prep = prepare_load(slot ^ 1);
      // ---- coalesced write-out: consecutive positions of one digit are consecutive output addresses
      K* __restrict__ kout = reinterpret_cast<K*>(a.keys_out);
      V* __restrict__ vout = reinterpret_cast<V*>(a.vals_out);
      {
        const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
        const bool two = a.tw_out != 0;
        const uint32_t ksrc = st_base + tid * (uint32_t)sizeof(K), vsrc = vst_base + tid * (uint32_t)sizeof(V);
        auto emit = [&](int j) {
          const uint32_t p = j * THREADS + tid;
          K k = lds_t<K>(ksrc + j * THREADS * (uint32_t)sizeof(K));
          const uint32_t o = lds_u32(go_base + digit_of<K>(k, shift, mask) * 4u) + p;
          if (two) k = tw_apply_out<K>(k, sg, fl, fp);
          st_global<K>(kout, o, k);
        };
        if (full) {
  #pragma unroll
          for (int j = 0; j < IPT; ++j) emit(j);
        } else {
  #pragma unroll
          for (int j = 0; j < IPT; ++j) {
            const uint32_t p = j * THREADS + tid;
            if (p < cnt) {
              K k = lds_t<K>(ksrc + j * THREADS * (uint32_t)sizeof(K));
              const uint32_t o = lds_u32(go_base + digit_of<K>(k, shift, mask) * 4u) + p;
              if (two) k = tw_apply_out<K>(k, sg, fl, fp);
              st_global<K>(kout, o, k);
              if (VB) st_global<V>(vout, o, lds_t<V>(vsrc + j * THREADS * (uint32_t)sizeof(V)));
            }
          }
        }
      }
      if (tid < RADIX) prepare_finish(slot ^ 1, prep);
    } else {
      // ---- keys: shared memory (TMA-staged) -> registers
      mbar_wait(&sm.bar[slot], (it >> 1) & 1u);
      K key[IPT]; uint32_t pos[IPT];
      {
        const K* __restrict__ src = st + g.skew + tid;
        if (full) {
  #pragma unroll
          for (int j = 0; j < IPT; ++j) key[j] = src[j * THREADS];
        } else {
  #pragma unroll
          for (int j = 0; j < IPT; ++j) key[j] = (tid + j * THREADS < cnt) ? src[j * THREADS] : (K)~(K)0;
        }
      }
      if (a.tw_in) {
        const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
  #pragma unroll
        for (int j = 0; j < IPT; ++j) key[j] = tw_apply_in<K>(key[j], sg, fl, fp);
      }
      // ---- rank: the atomicAdd returns the key's final slot in the reorder buffer
      if (!sm.skewed[slot]) {
        if (full) {
  #pragma unroll
          for (int j = 0; j < IPT; ++j) pos[j] = atomicAdd(&ctr[digit_of<K>(key[j], shift, mask)], 1u);
        } else {
  #pragma unroll
          for (int j = 0; j < IPT; ++j)
            if (tid + j * THREADS < cnt) pos[j] = atomicAdd(&ctr[digit_of<K>(key[j], shift, mask)], 1u);
        }
      } else {
        unsigned hot_d = sm.hot[slot];
        const unsigned lt = (1u << lane) - 1u;
  #pragma unroll
        for (int j = 0; j < IPT; ++j) {
          const bool v = tid + j * THREADS < cnt;
          const unsigned d = digit_of<K>(key[j], shift, mask);
          const unsigned hot = __ballot_sync(0xffffffffu, v && d == hot_d);
          if (v && d == hot_d) {
            unsigned b = 0;
            if ((hot & lt) == 0u) b = atomicAdd(&ctr[d], (uint32_t)__popc(hot));
            pos[j] = __shfl_sync(hot, b, __ffs(hot) - 1) + (uint32_t)__popc(hot & lt);
          } else if (v) {
            pos[j] = atomicAdd(&ctr[d], 1u);
          }
        }
      }
      V val[VB ? IPT : 1];
      if (VB) {
        const V* __restrict__ vsrc = vst + g.vskew + tid;
  #pragma unroll
        for (int j = 0; j < IPT; ++j)
          if (full || tid + j * THREADS < cnt) val[j] = vsrc[j * THREADS];
      }
      __syncthreads();          // every thread has its keys (and values) in registers: the staging buffers become the reorder buffers
  #pragma unroll
      for (int j = 0; j < IPT; ++j)
        if (full || tid + j * THREADS < cnt) { st[pos[j]] = key[j]; if (VB) vst[pos[j]] = val[j]; }
      __syncthreads();          // reorder complete; geom[slot ^ 1] (written by the producer above) is visible
      // ---- digit owners fetch the next tile's counts and destinations; the loads complete behind the write-out
      Prep prep{}; if (tid < RADIX) prep = prepare_load(slot ^ 1);
      // ---- coalesced write-out: consecutive positions of one digit are consecutive output addresses
      const uint32_t* __restrict__ go = sm.goff[slot];
      K* __restrict__ kout = reinterpret_cast<K*>(a.keys_out);
      V* __restrict__ vout = reinterpret_cast<V*>(a.vals_out);
      if (a.tw_out) {
        const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
  #pragma unroll
        for (int j = 0; j < IPT; ++j) {
          const uint32_t p = j * THREADS + tid;
          if (full || p < cnt) {
            const K k = st[p];
            const uint32_t d = digit_of<K>(k, shift, mask);
            const uint32_t o = go[d] + p;
            st_global<K>(kout, o, tw_apply_out<K>(k, sg, fl, fp));
            if (VB) st_global<V>(vout, o, vst[p]);
          }
        }
      } else {
  #pragma unroll
        for (int j = 0; j < IPT; ++j) {
          const uint32_t p = j * THREADS + tid;
          if (full || p < cnt) {
            const K k = st[p];
            const uint32_t d = digit_of<K>(k, shift, mask);
            const uint32_t o = go[d] + p;
            st_global<K>(kout, o, k);
            if (VB) st_global<V>(vout, o, vst[p]);
          }
        }
      }
      if (tid < RADIX) prepare_finish(slot ^ 1, prep);
    }
    __syncthreads();          // the slot (and cnt / kptr of this slot) may be overwritten from here on
    if (tid == PRODUCER) { tk_a = tk_b; td_a = td_b; }
  }
}

// ===============================================================================================================
// scatter_stable_fast_kernel -- the stable segment scatter (MODE_SEG, match-mask ranking) on the same one-tile-ahead
// preparation as scatter_fast_kernel: digit starts come from the scan of the tile's known counts (so no post-rank block scan
// and no padding correction), destinations are already in shared memory when the tile starts, values are read before the
// first barrier so keys and values are reordered in one phase: four block barriers per tile instead of seven.
// ===============================================================================================================
template <typename K, int VB, int THREADS, int IPT, bool PEER = false>
struct StableFastSmem {
  static constexpr int TILE = THREADS * IPT;
  static constexpr int WARPS = THREADS / 32;
  using V = typename ValType<VB>::type;
  static constexpr int SLACK = 16 / sizeof(K), VSLACK = 16 / sizeof(V);
  alignas(16) K stage[2][TILE + SLACK];
  alignas(16) V vstage[VB ? 2 : 1][VB ? TILE + VSLACK : 1];
  uint64_t dstk[PEER ? MAX_PARTS : 1], dstv[PEER ? MAX_PARTS : 1];   // PEER: base addresses of the destination buffers
  uint8_t gdst[PEER ? 2 : 1][PEER ? RADIX : 4];                      // PEER: per digit, which destination buffer
  uint32_t bsum[PEER ? RADIX : 1];                                   // PEER (inside prepare only): per bucket, keys of it in this source's earlier tiles
  uint16_t bexcl[PEER ? RADIX + 1 : 2];                              // PEER (inside prepare only): per bucket, where its keys start inside the tile
  // PEER, at most 32 exchange buckets: the write-out walks DESTINATION-aligned groups of 32 elements, so every warp store covers one
  // 128-byte line of the peer's buffer (a store that straddles two lines crosses NVLink as two partial packets: 400 vs 670 GB/s,
  // profiles/r02_ubench_peer.jsonl modes 4 / 2).  Per bucket: {first position - misalignment, start | end << 16, goff, destination};
  // grp[b] = number of groups of the buckets before b, grp[32] = all groups of the tile.
  alignas(16) uint4 brec[PEER ? 2 : 1][PEER ? 32 : 1];
  uint32_t grp[PEER ? 2 : 1][PEER ? 33 : 1];
  uint32_t goff[2][RADIX];          // per digit: (global start - start inside the tile) mod 2^32; output index = goff[d] + position
  alignas(16) uint32_t match[2][WARPS * RADIX];   // per-warp match masks, two alternating sets (always zero between rows)
  alignas(16) uint16_t wcnt[WARPS * RADIX];       // per-warp counters, later per-warp start positions; zero at tile start
  uint32_t excl[2][RADIX];                        // tile-local exclusive start of every digit (scan of the tile's counts)
  uint32_t scratch[2][8];
  alignas(8) uint64_t bar[2];
  TileGeom geom[2];
};

template <typename K, int VB, int THREADS, int IPT, int OCC, bool PEER = false>
__global__ void __launch_bounds__(THREADS, OCC) scatter_stable_fast_kernel(const __grid_constant__ ScatterArgs a) {
  pdl_wait();
  using SM = StableFastSmem<K, VB, THREADS, IPT, PEER>;
  using V = typename SM::V;
  constexpr int TILE = SM::TILE, WARPS = SM::WARPS;
  constexpr unsigned PRODUCER = THREADS - 1;
  static_assert(THREADS >= 2 * RADIX, "the digit owners (warps 0-7) must not include the producer's warp");
  static_assert(TILE < 65536, "per-warp start positions are 16-bit");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const K* __restrict__ keys_in = reinterpret_cast<const K*>(a.keys_in);
  const V* __restrict__ vals_in = reinterpret_cast<const V*>(a.vals_in);
  const uint32_t num_tiles = *a.num_tiles_ptr;
  const int shift = a.shift; const uint32_t mask = a.mask;

  auto stage_tile = [&](int slot, uint32_t t, const TileDesc& td) {
    TileGeom g;
    g.tile = t; g.skew = 0; g.vskew = 0;
    if (t < num_tiles) {
      g.off = td.off; g.cnt = td.cnt; g.seg = td.seg; g.tile_in_seg = td.tile_in_seg;
      const BulkWindow<K> bw(keys_in, g.off, g.cnt);
      g.skew = bw.skew;
      uint32_t bytes = bw.bytes;
      fence_proxy_async();
      if (VB) {
        const BulkWindow<V> vw(vals_in, g.off, g.cnt);
        g.vskew = vw.skew;
        bytes += vw.bytes;
        mbar_expect_tx(&sm.bar[slot], bytes);
        bulk_g2s(&sm.vstage[VB ? slot : 0][0], vw.src, vw.bytes, &sm.bar[slot]);
      } else {
        mbar_expect_tx(&sm.bar[slot], bytes);
      }
      bulk_g2s(&sm.stage[slot][0], bw.src, bw.bytes, &sm.bar[slot]);
    } else {
      g.off = 0; g.cnt = 0; g.seg = 0; g.tile_in_seg = 0;
    }
    sm.geom[slot] = g;
  };
  // digit owners (threads 0..255), in two steps like scatter_fast_kernel: the global loads, then (behind the owner's share of the
  // write-out) the scan and the shared-memory tables
  struct Prep { uint32_t c; uint64_t gstart; bool live; };
  auto prepare_load = [&](int slot) -> Prep {
    Prep p; p.c = 0; p.gstart = 0;
    const TileGeom g = sm.geom[slot];
    p.live = g.tile < num_tiles;
    if (!p.live) return p;
    p.c = a.tile_cnt[(uint64_t)g.tile * RADIX + tid];
    const uint32_t grp = g.tile / HIST_GROUP;
    p.gstart = (PEER ? 0ull : a.bins[(uint64_t)g.seg * RADIX + tid]) + a.tile_off[(uint64_t)g.tile * RADIX + tid];
    if (g.tile - g.tile_in_seg < grp * HIST_GROUP) p.gstart += a.carry[(uint64_t)grp * RADIX + tid];
    return p;
  };
  auto prepare_finish = [&](int slot, const Prep& p) {
    if (!p.live) return;
    const uint32_t c = p.c; const uint64_t gstart = p.gstart;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) sm.scratch[slot][w] = inc;
    if (PEER) sm.bsum[PEER ? tid : 0] = 0;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    uint32_t woff = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) woff += ((unsigned)j < w) ? sm.scratch[slot][j] : 0u;
    const uint32_t excl = woff + inc - c;
    sm.excl[slot][tid] = excl;
    if (!PEER) {
      sm.goff[slot][tid] = (uint32_t)gstart - excl;          // n < 2^32: indices wrap correctly in 32 bits
    } else {
      // exchange: the digits of one bucket (digit >> xshift) leave the tile as ONE run, chunk after chunk in the bucket's region of
      // its destination buffer.  Start of this tile's chunk = bucket start + keys of the bucket in this source's earlier tiles.
      const uint32_t b = tid >> a.xshift;
      if (gstart) atomicAdd(&sm.bsum[PEER ? b : 0], (uint32_t)gstart);
      if ((tid & ((1u << a.xshift) - 1u)) == 0) sm.bexcl[PEER ? b : 0] = (uint16_t)excl;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t gb = (uint32_t)a.bins[b] + sm.bsum[PEER ? b : 0] - (uint32_t)sm.bexcl[PEER ? b : 0];
      sm.goff[slot][tid] = gb;
      sm.gdst[PEER ? slot : 0][PEER ? tid : 0] = a.digit_dest[b];
      const uint32_t nbk = (uint32_t)RADIX >> a.xshift;
      if (nbk <= 32u) {
        const TileGeom gg = sm.geom[slot];
        if (tid == 0) sm.bexcl[PEER ? nbk : 0] = (uint16_t)gg.cnt;            // (padding keys rank last: real keys end at cnt)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (w == 0) {                                                          // lane = bucket
          uint32_t ng = 0;
          if (lane < nbk) {
            const uint32_t s0 = sm.bexcl[PEER ? lane : 0];
            uint32_t e0 = sm.bexcl[PEER ? lane + 1 : 0];
            if (e0 > gg.cnt) e0 = gg.cnt;
            const uint32_t gl = sm.goff[slot][lane << a.xshift];
            const uint32_t mis = (gl + s0) & 31u;                              // destination index of the bucket's first element, mod 32
            ng = e0 > s0 ? (e0 - s0 + mis + 31u) >> 5 : 0u;
            sm.brec[PEER ? slot : 0][PEER ? lane : 0] = make_uint4(s0 - mis, s0 | (e0 << 16), gl, (uint32_t)a.digit_dest[lane]);
          }
          uint32_t inc = ng;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
          }
          sm.grp[PEER ? slot : 0][PEER ? lane : 0] = inc - ng;
          if (lane == 31) sm.grp[PEER ? slot : 0][PEER ? 32 : 0] = inc;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");          // bsum / bexcl are reused by the next prepare
    }
  };

  uint32_t tk_a = 0, tk_b = 0;
  TileDesc td_a{}, td_b{};
  if (tid == PRODUCER) {
    mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
    const uint32_t t0 = blockIdx.x;
    tk_a = t0 + gridDim.x;
    TileDesc td{};
    if (t0 < num_tiles) td = a.descs[t0];
    if (tk_a < num_tiles) td_a = a.descs[tk_a];
    stage_tile(0, t0, td);
  }
  {
    uint4* z = reinterpret_cast<uint4*>(sm.match);
    for (int i = tid; i < 2 * WARPS * RADIX / 4; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
    uint4* zc = reinterpret_cast<uint4*>(sm.wcnt);
    for (int i = tid; i < WARPS * RADIX / 8; i += THREADS) zc[i] = make_uint4(0, 0, 0, 0);
  }
  if (PEER && (int)tid < a.num_parts && tid < (unsigned)MAX_PARTS) {
    sm.dstk[PEER ? tid : 0] = a.dst_keys[tid];
    if (VB) sm.dstv[PEER ? tid : 0] = a.dst_vals[tid];
  }
  __syncthreads();
  if (tid < RADIX) prepare_finish(0, prepare_load(0));
  __syncthreads();

  for (uint32_t it = 0;; ++it) {
    const int slot = (int)(it & 1u);
    const TileGeom g = sm.geom[slot];
    if (g.tile >= num_tiles) break;
    const uint32_t cnt = g.cnt;
    const bool full = cnt == (uint32_t)TILE;
    if (tid == PRODUCER) {
      stage_tile(slot ^ 1, tk_a, td_a);
      tk_b = tk_a + gridDim.x;
      if (tk_b < num_tiles) td_b = a.descs[tk_b];
    }
    K* __restrict__ st = &sm.stage[slot][0];
    V* __restrict__ vst = &sm.vstage[VB ? slot : 0][0];

    // ---- keys (and values): shared memory (TMA-staged) -> registers, warp-contiguous layout (row j of warp w = keys w*32*IPT + j*32 ..)
    mbar_wait(&sm.bar[slot], (it >> 1) & 1u);
    K key[IPT]; uint32_t pos[IPT]; V val[VB ? IPT : 1];
    const uint32_t ibase = w * (32u * IPT) + lane;
    {
      const K* __restrict__ src = st + g.skew + ibase;
      const V* __restrict__ vsrc = vst + g.vskew + ibase;
      if (full) {
#pragma unroll
        for (int j = 0; j < IPT; ++j) { key[j] = src[j * 32]; if (VB) val[j] = vsrc[j * 32]; }
      } else {
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
          const bool v = ibase + j * 32 < cnt;
          key[j] = v ? src[j * 32] : (K)~(K)0;            // padding keys (all ones) sit last and rank last
          if (VB && v) val[j] = vsrc[j * 32];
        }
      }
    }
    if (a.tw_in) {
      const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
#pragma unroll
      for (int j = 0; j < IPT; ++j)
        if (full || ibase + j * 32 < cnt) key[j] = tw_apply_in<K>(key[j], sg, fl, fp);
    }
    // ---- stable ranking inside the warp's share (scatter.cuh, scatter_tile)
    {
      uint16_t* wc = sm.wcnt + w * RADIX;
      const unsigned lt = (1u << lane) - 1u, lbit = 1u << lane;
      // the lanes on the row's hot digit (the most frequent digit of the warp's previous row) meet through one vote instead of
      // same-address atomics, which serialise on skewed digits (local_sort.cuh, lsd_sort_item)
      unsigned hot_d = 0;
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        uint32_t* wm = sm.match[j & 1] + w * RADIX;
        const unsigned d = digit_of<K>(key[j], shift, mask);
        const bool is_hot = d == hot_d;
        const unsigned hot = __ballot_sync(0xffffffffu, is_hot);
        if (!is_hot) atomicOr(&wm[d], lbit);
        __syncwarp();
        const unsigned peers = is_hot ? hot : wm[d];
        __syncwarp();
        const unsigned below = __popc(peers & lt);
        unsigned b = 0;
        if (below == 0) { b = wc[d]; wc[d] = (uint16_t)(b + __popc(peers)); if (!is_hot) wm[d] = 0; }
        b = __shfl_sync(0xffffffffu, b, __ffs(peers) - 1);
        pos[j] = b + below;
        hot_d = __reduce_max_sync(0xffffffffu, ((unsigned)__popc(peers) << 8) | d) & 0xFFu;
      }
    }
    __syncthreads();          // (1) all per-warp counts are final; every thread holds its keys and values in registers
    if (tid < RADIX) {        // per-warp start positions: the digit's start (known a tile ahead) + counts of the lower warps
      uint32_t run = sm.excl[slot][tid];
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) {
        const uint32_t c = sm.wcnt[ww * RADIX + tid];
        sm.wcnt[ww * RADIX + tid] = (uint16_t)run;
        run += c;
      }
    }
    __syncthreads();          // (2)
    {
      const uint16_t* wc = sm.wcnt + w * RADIX;
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        const uint32_t q = pos[j] + wc[digit_of<K>(key[j], shift, mask)];
        if (full || q < cnt) { st[q] = key[j]; if (VB) vst[q] = val[j]; }      // padding ranks after every real key: q >= cnt
      }
    }
    __syncthreads();          // (3) reorder complete; geom[slot ^ 1] is visible
    Prep prep{}; if (tid < RADIX) prep = prepare_load(slot ^ 1);      // (the loads complete behind the write-out)
    {                         // the per-warp counters are free again: zero them for the next tile
      uint4* zc = reinterpret_cast<uint4*>(sm.wcnt);
      for (int i = tid; i < WARPS * RADIX / 8; i += THREADS) zc[i] = make_uint4(0, 0, 0, 0);
    }
    const uint32_t* __restrict__ go = sm.goff[slot];
    K* __restrict__ kout = reinterpret_cast<K*>(a.keys_out);
    V* __restrict__ vout = reinterpret_cast<V*>(a.vals_out);
    if (PEER && ((uint32_t)RADIX >> a.xshift) <= 32u) {
      // destination-aligned write-out: group g of the tile = 32 consecutive destination slots of one bucket, one warp store per array
      const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
      const bool two = a.tw_out != 0;
      const uint32_t nbk = (uint32_t)RADIX >> a.xshift;
      const uint32_t my_g = lane < nbk ? sm.grp[PEER ? slot : 0][PEER ? lane : 0] : 0xFFFFFFFFu;      // groups before bucket `lane`
      const uint32_t ngroups = sm.grp[PEER ? slot : 0][PEER ? 32 : 0];
      for (uint32_t gidx = w; gidx < ngroups; gidx += WARPS) {
        const uint32_t b = (uint32_t)__popc(__ballot_sync(0xffffffffu, my_g <= gidx)) - 1u;           // the bucket of this group (empty buckets share a start: the last wins)
        const uint4 r = sm.brec[PEER ? slot : 0][PEER ? b : 0];
        const uint32_t p = r.x + ((gidx - __shfl_sync(0xffffffffu, my_g, b)) << 5) + lane;             // (r.x may be "negative": wraps back into range)
        if ((int32_t)p >= (int32_t)(r.y & 0xFFFFu) && p < (r.y >> 16)) {
          K k = st[p];
          if (two) k = tw_apply_out<K>(k, sg, fl, fp);
          const uint32_t o = r.z + p;
          st_global<K>(reinterpret_cast<K*>(sm.dstk[PEER ? r.w : 0]), o, k);
          if (VB) st_global<V>(reinterpret_cast<V*>(sm.dstv[PEER ? r.w : 0]), o, vst[p]);
        }
      }
    } else {
      const K sg = (K)a.tw.sign_mask, fl = (K)a.tw.float_mask, fp = (K)a.tw.flip_mask;
      const bool two = a.tw_out != 0;
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        const uint32_t p = j * THREADS + tid;
        if (full || p < cnt) {
          K k = st[p];
          const uint32_t d = digit_of<K>(k, shift, mask);
          if (two) k = tw_apply_out<K>(k, sg, fl, fp);
          if (PEER) {
            const uint32_t o = go[d] + p, dd = sm.gdst[PEER ? slot : 0][PEER ? d : 0];
            st_global<K>(reinterpret_cast<K*>(sm.dstk[PEER ? dd : 0]), o, k);
            if (VB) st_global<V>(reinterpret_cast<V*>(sm.dstv[PEER ? dd : 0]), o, vst[p]);
          } else {
            const uint32_t o = go[d] + p;
            st_global<K>(kout, o, k);
            if (VB) st_global<V>(vout, o, vst[p]);
          }
        }
      }
    }
    if (tid < RADIX) prepare_finish(slot ^ 1, prep);
    __syncthreads();          // (4)
    if (tid == PRODUCER) { tk_a = tk_b; td_a = td_b; }
  }
}

}  // namespace b200
