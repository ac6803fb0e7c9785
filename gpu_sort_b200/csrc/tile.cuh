// tile.cuh -- in-tile digit ranking: turns the digits of one tile of keys (held in registers) into tile-local
// destination positions, i.e. the permutation that groups the tile by digit.
//
// Two ranking modes (tools/ubench_rank.cu, profiles/ubench_rank_r01.jsonl measured both on B200):
//   ORDERED = false : one shared-memory atomicAdd-with-return per key on 256 block-shared counters.
//                     ~3200 G ranks/s full chip for uniform digits; order inside a digit is arbitrary, which is
//                     all the unstable MSB path needs (the reference's partition also uses one SMEM atomicAdd per
//                     key, msb/src/sort/cuda_radix_sort.h:129-133).
//   ORDERED = true  : stable ranking: 8 x VOTE digit matching inside the warp + warp-private running counters
//                     (cost independent of the key distribution, no same-address atomics).  Needed by every LSB
//                     pass (reference: BlockRadixRank, lsb/cub/cub/block/block_radix_rank.cuh:341-430).
#pragma once
#include "common.cuh"

namespace b200 {

template <int THREADS, bool ORDERED>
struct RankSmem {
  static constexpr int WARPS = THREADS / 32;
  uint32_t cnt[ORDERED ? WARPS * RADIX : RADIX];  // ORDERED: per-warp counters, later per-warp start positions
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
template <int THREADS, int IPT, bool ORDERED>
__device__ __forceinline__ void tile_positions(const uint32_t (&dg)[IPT], uint32_t valid, int rows, uint32_t pad,
                                               uint32_t pad_digit, uint32_t (&pos)[IPT],
                                               RankSmem<THREADS, ORDERED>& sm, uint32_t& my_total, uint32_t& my_excl) {
  constexpr int WARPS = THREADS / 32;
  const unsigned tid = threadIdx.x;
  if (ORDERED) {
    const unsigned lane = tid & 31u, w = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = tid; i < WARPS * RADIX; i += THREADS) sm.cnt[i] = 0;
    __syncthreads();
    uint32_t* wc = sm.cnt + w * RADIX;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      if (j < rows) {                         // block-uniform
        const unsigned d = dg[j];
        const unsigned peers = match_digit(d);
        const unsigned base = wc[d];
        __syncwarp();
        const unsigned below = __popc(peers & lt);
        if (below == 0) wc[d] = base + __popc(peers);
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
      if (j < rows) pos[j] += wc[dg[j]];
  } else {
    if (tid < RADIX) sm.cnt[tid] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if ((valid >> j) & 1u) pos[j] = atomicAdd(&sm.cnt[dg[j]], 1u);
    __syncthreads();
    const uint32_t total = tid < RADIX ? sm.cnt[tid] : 0u;
    const uint32_t excl = block_excl_scan_256(total, sm.scratch);
    if (tid < RADIX) sm.bin_start[tid] = excl;
    my_total = total; my_excl = excl;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if ((valid >> j) & 1u) pos[j] += sm.bin_start[dg[j]];
  }
}

}  // namespace b200
