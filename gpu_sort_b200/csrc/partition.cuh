// partition.cuh -- the bucket scatter: one read + one write sweep that splits every segment by one 8-bit digit.
//
// One kernel serves (a) every pass of the stable LSB sort (single segment, onesweep-style: digit starts from the
// up-front histogram + decoupled look-back for the tile prefix), (b) every level of the MSB sort (many segments;
// tiles described by a device-built TileDesc list) and (c) the multi-GPU send partition (range mode).  Replaces
// rdxsrt_partition_keys (msb/src/sort/cuda_radix_sort.h:363-479) and DeviceRadixSortDownsweepKernel
// (lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:164-196).
//
// Structure (persistent CTAs, tiles taken from an atomic ticket, which is what makes the look-back deadlock-free):
//   * the NEXT tile's keys are staged into shared memory by a TMA bulk copy (cp.async.bulk + mbarrier) issued by
//     one producer thread while the current tile is being ranked -> HBM reads overlap everything else and cost no
//     registers / issue slots;
//   * ranking in registers (tile.cuh);
//   * digit owners (threads 0..255) publish the tile aggregate and look back LOOKBACK_BATCH predecessors per L2
//     round trip;
//   * keys (and values) are reordered through shared memory (the consumed staging buffer is reused) so that every
//     digit's run leaves as consecutive addresses -> coalesced stores.
#pragma once
#include "async.cuh"
#include "tile.cuh"

namespace b200 {

struct PartArgs {
  const void* keys_in; void* keys_out;
  const void* vals_in; void* vals_out;
  const TileDesc* descs;          // segment mode: tile -> (offset, count, segment, tile in segment); nullptr otherwise
  const uint32_t* num_tiles_ptr;  // segment mode: device-side tile count
  uint32_t num_tiles;             // single-segment mode
  uint64_t base, n;               // single-segment mode: the launch covers keys [base, base+n)
  const uint64_t* bins;           // [segment][256] absolute output index of the start of each (segment, digit);
                                  // the unordered (MSB) instantiation advances these as chunk-reservation cursors
  uint64_t* bins_next;            // single-segment mode: last tile writes bins + portion counts here (or nullptr)
  uint32_t* status;               // [tile][256] look-back words, zeroed before the launch
  uint32_t* ticket;               // zeroed before the launch
  int shift; uint32_t mask;
  int tw_in, tw_out;
  Twiddle tw;
  // range mode (multi-GPU send partition): digit = #{ j < num_parts-1 : splitters[j] <= (key >> shift) }
  const uint32_t* splitters; int num_parts;
};

constexpr int MAX_PARTS = 16;
constexpr int LOOKBACK_BATCH = 8;

struct TileGeom {     // per staging slot, written by the producer thread
  uint64_t off;       // first key index
  uint32_t cnt, seg, tile_in_seg, tile, skew, pad;
};

template <typename K, int VB, int THREADS, int IPT, bool ORDERED>
struct PartSmem {
  static constexpr int TILE = THREADS * IPT;
  static constexpr int SLACK = 16 / sizeof(K);     // alignment skew of an unaligned tile start
  using V = typename ValType<VB>::type;
  alignas(16) K stage[2][TILE + SLACK];
  alignas(16) V vals[VB ? TILE : 1];
  RankSmem<THREADS, ORDERED> rank;
  uint64_t goff[RADIX];
  alignas(8) uint64_t bar[2];
  TileGeom geom[2];
  uint32_t split[MAX_PARTS];
};

// digit of a key under the launch's rule: plain bit field, or destination rank in range mode
template <typename K>
__device__ __forceinline__ uint32_t part_digit(K k, const PartArgs& a, const uint32_t* split) {
  if (a.splitters == nullptr) return digit_of<K>(k, a.shift, a.mask);
  const uint32_t b = (uint32_t)(k >> a.shift);
  uint32_t d = 0;
#pragma unroll
  for (int j = 0; j < MAX_PARTS - 1; ++j) d += (j < a.num_parts - 1 && b >= split[j]) ? 1u : 0u;
  return d;
}

template <typename K, int VB, int THREADS, int IPT, bool ORDERED>
__global__ void __launch_bounds__(THREADS, 2) partition_kernel(const __grid_constant__ PartArgs a) {
  using V = typename ValType<VB>::type;
  using SM = PartSmem<K, VB, THREADS, IPT, ORDERED>;
  constexpr int TILE = THREADS * IPT;
  constexpr unsigned PRODUCER = THREADS - 1;     // not a digit owner (THREADS > 256)
  static_assert(THREADS > RADIX, "producer thread must not own a digit");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const K* __restrict__ keys_in = reinterpret_cast<const K*>(a.keys_in);
  K* __restrict__ keys_out = reinterpret_cast<K*>(a.keys_out);
  const V* __restrict__ vals_in = reinterpret_cast<const V*>(a.vals_in);
  V* __restrict__ vals_out = reinterpret_cast<V*>(a.vals_out);
  const uint32_t num_tiles = a.descs ? *a.num_tiles_ptr : a.num_tiles;

  // producer: describe tile `t`, arm the slot's barrier and launch the bulk copy of its keys
  auto stage_tile = [&](int slot, uint32_t t, const TileDesc& td) {
    TileGeom g;
    g.tile = t; g.pad = 0;
    if (t < num_tiles) {
      if (a.descs) { g.off = td.off; g.cnt = td.cnt; g.seg = td.seg; g.tile_in_seg = td.tile_in_seg; }
      else {
        const uint64_t rel = (uint64_t)t * TILE;
        g.off = a.base + rel; g.cnt = (uint32_t)(a.n - rel < (uint64_t)TILE ? a.n - rel : (uint64_t)TILE);
        g.seg = 0; g.tile_in_seg = t;
      }
      const BulkWindow<K> bw(keys_in, g.off, g.cnt);
      g.skew = bw.skew;
      fence_proxy_async();
      mbar_expect_tx(&sm.bar[slot], bw.bytes);
      bulk_g2s(&sm.stage[slot][0], bw.src, bw.bytes, &sm.bar[slot]);
    } else {
      g.off = 0; g.cnt = 0; g.seg = 0; g.tile_in_seg = 0; g.skew = 0;
    }
    sm.geom[slot] = g;
  };

  // Tile sequence.  ORDERED (LSB / range mode): dynamic tickets -- a tile is only ever taken by a CTA that is about
  // to process it, so every predecessor the look-back may wait for is already running; the ticket for the next
  // tile is requested at the top of an iteration and consumed (TMA issue) after the ranking.  Unordered (MSB):
  // tiles are independent (output chunks are reserved with one global atomicAdd per digit, as the reference does,
  // cuda_radix_sort.h:408-417), so they are dealt round-robin and the next descriptor is fetched a full iteration ahead.
  uint32_t tk_next = 0;      // producer only: tile for the next iteration
  if (tid == PRODUCER) {
    mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
    const uint32_t t0 = ORDERED ? atomicAdd(a.ticket, 1u) : blockIdx.x;
    TileDesc td{};
    if (a.descs && t0 < num_tiles) td = a.descs[t0];
    stage_tile(0, t0, td);
  }
  if (a.splitters != nullptr && tid < MAX_PARTS) sm.split[tid] = (int)tid < a.num_parts - 1 ? a.splitters[tid] : 0xFFFFFFFFu;
  __syncthreads();

  for (uint32_t it = 0;; ++it) {
    const int slot = (int)(it & 1u);
    const TileGeom g = sm.geom[slot];
    const uint32_t tile = g.tile;
    if (tile >= num_tiles) break;
    const uint32_t cnt = g.cnt;
    const bool first = g.tile_in_seg == 0;
    const uint32_t first_tile = tile - g.tile_in_seg;
    const uint64_t* bins = a.bins + (uint64_t)g.seg * RADIX;

    // producer: start fetching what the next step needs (results are consumed after the ranking)
    TileDesc td_next{};
    if (tid == PRODUCER) {
      tk_next = ORDERED ? atomicAdd(a.ticket, 1u) : tile + gridDim.x;
      if (!ORDERED && a.descs && tk_next < num_tiles) td_next = a.descs[tk_next];
    }

    // ---- keys of this tile: shared memory (TMA-staged) -> registers
    mbar_wait(&sm.bar[slot], (it >> 1) & 1u);
    K* __restrict__ st = &sm.stage[slot][0];
    K key[IPT]; uint32_t pos[IPT];
    uint32_t valid = 0;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const uint32_t idx = ORDERED ? (w * (32 * IPT) + j * 32 + lane) : (j * THREADS + tid);
      K k = (K)~(K)0;
      if (idx < cnt) {
        k = st[g.skew + idx];
        if (a.tw_in) k = twiddle_in<K>(k, a.tw);
        valid |= 1u << j;
      }
      key[j] = k;
    }
    auto dfn = [&](K k) { return part_digit<K>(k, a, sm.split); };

    // ---- rank inside the tile
    uint32_t my_total, my_excl;
    tile_positions<THREADS, IPT, ORDERED>(key, dfn, valid, IPT, (uint32_t)TILE - cnt, a.mask, pos, sm.rank, my_total, my_excl);

    // ---- producer: the other slot is free (its tile finished last iteration) -> prefetch the next tile into it
    if (tid == PRODUCER) {
      if (ORDERED && a.descs && tk_next < num_tiles) td_next = a.descs[tk_next];
      stage_tile(slot ^ 1, tk_next, td_next);
    }

    // ---- digit owners: publish the tile aggregate, look back for the exclusive prefix, derive global offsets
    if (tid < RADIX) {
      uint64_t gstart;
      if (!ORDERED) {
        // unordered placement: reserve this tile's chunk of the (segment, digit) sub-bucket
        gstart = my_total ? atomicAdd(const_cast<unsigned long long*>(reinterpret_cast<const unsigned long long*>(bins)) + tid,
                                      (unsigned long long)my_total) : 0ull;
      } else {
      uint32_t* stw = a.status + (uint64_t)tile * RADIX + tid;
      uint32_t excl_g = 0;
      if (first) {
        st_status(stw, ST_PREFIX | my_total);
      } else {
        st_status(stw, ST_AGG | my_total);
        // Decoupled look-back, LOOKBACK_BATCH predecessors per round trip: the loads of a batch are independent, so
        // a walk of depth D costs ~D/BATCH L2 latencies instead of D.  A stale word is still valid (an aggregate
        // never changes, it is only upgraded to a prefix), so only not-yet-published words are re-read.
        int64_t t = (int64_t)tile - 1;
        bool done = false;
        while (!done) {
          uint32_t s[LOOKBACK_BATCH];
#pragma unroll
          for (int j = 0; j < LOOKBACK_BATCH; ++j)
            s[j] = (t - j >= (int64_t)first_tile) ? ld_status(a.status + (uint64_t)(t - j) * RADIX + tid) : ST_PREFIX;
#pragma unroll
          for (int j = 0; j < LOOKBACK_BATCH; ++j) {
            if (!done) {
              uint32_t v = s[j];
              while ((v >> 30) == 0) { __nanosleep(32); v = ld_status(a.status + (uint64_t)(t - j) * RADIX + tid); }
              excl_g += v & ST_VALUE_MASK;
              done = (v & ST_PREFIX) != 0;
            }
          }
          t -= LOOKBACK_BATCH;
        }
        st_status(stw, ST_PREFIX | (excl_g + my_total));
      }
      gstart = bins[tid] + excl_g;          // where this tile's run of digit `tid` begins
      if (a.bins_next != nullptr && tile == num_tiles - 1) a.bins_next[tid] = gstart + my_total;
      }
      sm.goff[tid] = gstart - my_excl;                     // output index = goff[digit] + position in tile
    }

    // ---- reorder through shared memory (every thread has its keys in registers: the staging buffer is reused)
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if ((valid >> j) & 1u) st[pos[j]] = key[j];
    if (VB) {
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        const uint32_t idx = ORDERED ? (w * (32 * IPT) + j * 32 + lane) : (j * THREADS + tid);
        if ((valid >> j) & 1u) sm.vals[pos[j]] = vals_in[g.off + idx];
      }
    }
    __syncthreads();

    // ---- coalesced write-out: consecutive positions of one digit are consecutive output addresses
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const uint32_t p = j * THREADS + tid;
      if (p < cnt) {
        K k = st[p];
        const uint64_t o = sm.goff[part_digit<K>(k, a, sm.split)] + p;
        if (a.tw_out) k = twiddle_out<K>(k, a.tw);
        keys_out[o] = k;
        if (VB) vals_out[o] = sm.vals[p];
      }
    }
    __syncthreads();   // the slot (and sm.vals / sm.goff) may be overwritten from here on
  }
}

}  // namespace b200
