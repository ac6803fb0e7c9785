// tile.cuh -- in-tile digit ranking: turns the digits of one tile of keys (held in registers) into tile-local
// destination positions, i.e. the permutation that groups the tile by digit.
//
// Two ranking modes (tools/ubench_rank.cu, profiles/ubench_rank_r01.jsonl measured the candidates on B200):
//   ORDERED = false : one shared-memory atomicAdd-with-return per key on 256 block-shared counters.
//                     ~3200 G ranks/s full chip for uniform digits; order inside a digit is arbitrary, which is
//                     all the unstable MSB path needs (the reference's partition also uses one SMEM atomicAdd per
//                     key, msb/src/sort/cuda_radix_sort.h:129-133).  A warp whose 32 digits are all equal
//                     (constant / heavily skewed input) adds once instead of serialising 32 same-address atomics.
//   ORDERED = true  : stable ranking: the lanes of a warp that hold the same digit find each other through a
//                     shared-memory atomicOr of their lane bit into a per-warp match mask (~2x the throughput of
//                     8 x VOTE matching, 4.7x MATCH.ANY on uniform digits), then rank against warp-private running
//                     counters.  Needed by every LSB pass (reference: BlockRadixRank,
//                     lsb/cub/cub/block/block_radix_rank.cuh:341-430).
#pragma once
#include "common.cuh"

namespace b200 {

template <int THREADS, bool ORDERED>
struct RankSmem {
  static constexpr int WARPS = THREADS / 32;
  uint32_t cnt[ORDERED ? WARPS * RADIX : RADIX];  // ORDERED: per-warp counters, later per-warp start positions
  uint32_t mask[ORDERED ? WARPS * RADIX : 1];     // ORDERED: per-warp match masks (all zero between rows)
  uint32_t bin_start[RADIX];                      // tile-local exclusive start of each digit
  uint32_t scratch[8];
};

// Item layout the caller must use:
//   ORDERED : item j of lane l of warp w is sequence element  w*rows*32 + j*32 + l   (warp-contiguous)
//   !ORDERED: any; only `valid` matters.
// `rows` (block-uniform, <= IPT) = items per thread in use.  valid bit j = item j holds a real key.
// In ORDERED mode invalid items MUST carry digit `pad_digit` (the largest digit value) and sit at the very end
// of the sequence; `pad` = number of such items in the tile (they are ranked after every real key and are
// subtracted from the digit's count).
// On return pos[j] = tile-local destination of item j; for threads < 256, my_total / my_excl = count and
// exclusive start of digit threadIdx.x.  Ends with a __syncthreads (sm.bin_start is readable).
template <int THREADS, int IPT, bool ORDERED, typename K, typename DigitFn>
__device__ __forceinline__ void tile_positions(const K (&key)[IPT], DigitFn dfn, uint32_t valid, int rows, uint32_t pad,
                                               uint32_t pad_digit, uint32_t (&pos)[IPT],
                                               RankSmem<THREADS, ORDERED>& sm, uint32_t& my_total, uint32_t& my_excl) {
  constexpr int WARPS = THREADS / 32;
  const unsigned tid = threadIdx.x;
  if (ORDERED) {
    const unsigned lane = tid & 31u, w = tid >> 5;
    const unsigned lt = (1u << lane) - 1u, lbit = 1u << lane;
    for (int i = tid; i < WARPS * RADIX; i += THREADS) { sm.cnt[i] = 0; sm.mask[i] = 0; }
    __syncthreads();
    uint32_t* wc = sm.cnt + w * RADIX;
    uint32_t* wm = sm.mask + w * RADIX;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      if (j < rows) {                         // block-uniform
        const unsigned d = dfn(key[j]);
        atomicOr(&wm[d], lbit);
        __syncwarp();
        const unsigned peers = wm[d];
        const unsigned base = wc[d];
        __syncwarp();
        const unsigned below = __popc(peers & lt);
        if (below == 0) { wc[d] = base + __popc(peers); wm[d] = 0; }
        __syncwarp();
        pos[j] = base + below;
      }
    }
    __syncthreads();
    uint32_t total = 0;
    if (tid < RADIX) {
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) total += sm.cnt[ww * RADIX + tid];
      if (tid == pad_digit) total -= pad;
    }
    const uint32_t excl = block_excl_scan_256(tid < RADIX ? total : 0u, sm.scratch);
    if (tid < RADIX) {
      sm.bin_start[tid] = excl;
      uint32_t run = excl;
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) {
        const uint32_t c = sm.cnt[ww * RADIX + tid];
        sm.cnt[ww * RADIX + tid] = run;
        run += c;
      }
    }
    my_total = total; my_excl = excl;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if (j < rows) pos[j] += wc[dfn(key[j])];
  } else {
    if (tid < RADIX) sm.cnt[tid] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const bool v = (valid >> j) & 1u;
      const unsigned d = dfn(key[j]);
      // warp-uniform digit (constant / heavily skewed input): rank = lane order, one add for the whole warp
      const unsigned d0 = __shfl_sync(0xffffffffu, d, 0);
      if (__all_sync(0xffffffffu, v && d == d0)) {
        unsigned base = 0;
        if ((tid & 31u) == 0) base = atomicAdd(&sm.cnt[d0], 32u);
        pos[j] = __shfl_sync(0xffffffffu, base, 0) + (tid & 31u);
      } else if (v) {
        pos[j] = atomicAdd(&sm.cnt[d], 1u);
      }
    }
    __syncthreads();
    const uint32_t total = tid < RADIX ? sm.cnt[tid] : 0u;
    const uint32_t excl = block_excl_scan_256(total, sm.scratch);
    if (tid < RADIX) sm.bin_start[tid] = excl;
    my_total = total; my_excl = excl;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if ((valid >> j) & 1u) pos[j] += sm.bin_start[dfn(key[j])];
  }
}

}  // namespace b200
