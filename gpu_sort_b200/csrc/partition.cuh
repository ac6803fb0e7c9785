// partition.cuh -- the bucket scatter: one read + one write sweep that splits every segment by one 8-bit digit.
//
// One kernel serves (a) every pass of the stable LSB sort (single segment, onesweep-style: digit starts from the
// up-front histogram + decoupled look-back for the tile prefix) and (b) every level of the MSB sort (many
// segments; tiles described by a device-built TileDesc list).  Replaces rdxsrt_partition_keys
// (msb/src/sort/cuda_radix_sort.h:363-479) and DeviceRadixSortDownsweepKernel
// (lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:164-196).
//
// Per tile: rank in registers (tile.cuh) -> reorder through shared memory so that every digit's run leaves as
// consecutive addresses -> coalesced stores.  Persistent CTAs take tiles from an atomic ticket, which is what makes
// the look-back deadlock-free (a tile only waits on lower tickets, all of which are already running).
#pragma once
#include "tile.cuh"

namespace b200 {

struct PartArgs {
  const void* keys_in; void* keys_out;
  const void* vals_in; void* vals_out;
  const Seg* segs;                // nullptr => single segment [base, base+n)
  const TileDesc* descs;          // segment mode: tile -> (segment, tile in segment)
  const uint32_t* num_tiles_ptr;  // segment mode: device-side tile count
  uint32_t num_tiles;             // single-segment mode
  uint64_t base, n;               // single-segment mode
  const uint64_t* bins;           // [segment][256] absolute output index of the start of each (segment, digit)
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

template <typename K, int VB, int THREADS, int IPT, bool ORDERED>
struct PartSmem {
  static constexpr int TILE = THREADS * IPT;
  using V = typename ValType<VB>::type;
  alignas(16) K keys[TILE];
  alignas(16) V vals[VB ? TILE : 1];
  RankSmem<THREADS, ORDERED> rank;
  uint64_t goff[RADIX];
  uint32_t split[MAX_PARTS];
  uint32_t tile;
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
__global__ void __launch_bounds__(THREADS) partition_kernel(const __grid_constant__ PartArgs a) {
  using V = typename ValType<VB>::type;
  using SM = PartSmem<K, VB, THREADS, IPT, ORDERED>;
  constexpr int TILE = THREADS * IPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const K* __restrict__ keys_in = reinterpret_cast<const K*>(a.keys_in);
  K* __restrict__ keys_out = reinterpret_cast<K*>(a.keys_out);
  const V* __restrict__ vals_in = reinterpret_cast<const V*>(a.vals_in);
  V* __restrict__ vals_out = reinterpret_cast<V*>(a.vals_out);
  const uint32_t num_tiles = a.segs ? *a.num_tiles_ptr : a.num_tiles;
  if (a.splitters != nullptr && tid < MAX_PARTS) sm.split[tid] = (int)tid < a.num_parts - 1 ? a.splitters[tid] : 0xFFFFFFFFu;

  while (true) {
    if (tid == 0) sm.tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t tile = sm.tile;
    if (tile >= num_tiles) break;

    // ---- tile geometry
    uint64_t tile_off; uint32_t cnt; bool first; const uint64_t* bins;
    if (a.segs) {
      const TileDesc td = a.descs[tile];
      const Seg sg = a.segs[td.seg];
      const uint64_t rel = (uint64_t)td.tile_in_seg * TILE;
      tile_off = sg.off + rel;
      cnt = (uint32_t)min((uint64_t)TILE, sg.cnt - rel);
      first = td.tile_in_seg == 0;
      bins = a.bins + (uint64_t)td.seg * RADIX;
    } else {
      const uint64_t rel = (uint64_t)tile * TILE;
      tile_off = a.base + rel;
      cnt = (uint32_t)min((uint64_t)TILE, a.n - rel);
      first = tile == 0;
      bins = a.bins;
    }

    // ---- load + digits
    K key[IPT]; uint32_t dg[IPT], pos[IPT];
    uint32_t valid = 0;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const uint32_t idx = ORDERED ? (w * (32 * IPT) + j * 32 + lane) : (j * THREADS + tid);
      K k = (K)~(K)0;
      if (idx < cnt) {
        k = keys_in[tile_off + idx];
        if (a.tw_in) k = twiddle_in<K>(k, a.tw);
        valid |= 1u << j;
      }
      key[j] = k;
      dg[j] = part_digit<K>(k, a, sm.split);
    }

    // ---- rank inside the tile
    uint32_t my_total, my_excl;
    tile_positions<THREADS, IPT, ORDERED>(dg, valid, IPT, (uint32_t)TILE - cnt, a.mask, pos, sm.rank, my_total, my_excl);

    // ---- digit owners: publish the tile aggregate, look back for the exclusive prefix, derive global offsets
    if (tid < RADIX) {
      uint32_t* st = a.status + (uint64_t)tile * RADIX + tid;
      uint32_t excl_g = 0;
      if (first) {
        st_status(st, ST_PREFIX | my_total);
      } else {
        st_status(st, ST_AGG | my_total);
        const uint32_t* p = st - RADIX;
        while (true) {
          uint32_t s = ld_status(p);
          while ((s >> 30) == 0) { __nanosleep(20); s = ld_status(p); }
          excl_g += s & ST_VALUE_MASK;
          if (s & ST_PREFIX) break;
          p -= RADIX;
        }
        st_status(st, ST_PREFIX | (excl_g + my_total));
      }
      const uint64_t gstart = bins[tid] + excl_g;          // where this tile's run of digit `tid` begins
      sm.goff[tid] = gstart - my_excl;                     // output index = goff[digit] + position in tile
      if (a.bins_next != nullptr && tile == num_tiles - 1) a.bins_next[tid] = gstart + my_total;
    }

    // ---- reorder through shared memory
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if ((valid >> j) & 1u) sm.keys[pos[j]] = key[j];
    if (VB) {
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        const uint32_t idx = ORDERED ? (w * (32 * IPT) + j * 32 + lane) : (j * THREADS + tid);
        if ((valid >> j) & 1u) sm.vals[pos[j]] = vals_in[tile_off + idx];
      }
    }
    __syncthreads();

    // ---- coalesced write-out: consecutive positions of one digit are consecutive output addresses
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const uint32_t p = j * THREADS + tid;
      if (p < cnt) {
        K k = sm.keys[p];
        const uint64_t o = sm.goff[part_digit<K>(k, a, sm.split)] + p;
        if (a.tw_out) k = twiddle_out<K>(k, a.tw);
        keys_out[o] = k;
        if (VB) vals_out[o] = sm.vals[p];
      }
    }
    __syncthreads();   // shared memory is reused by the next tile
  }
}

}  // namespace b200
