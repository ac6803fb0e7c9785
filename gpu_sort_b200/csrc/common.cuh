// common.cuh -- shared device-side building blocks of the B200 radix-sort library.
//
// Everything here is integer/byte work bounded by HBM bandwidth and the SM's shared-memory pipe; no tensor
// cores are involved (nothing on the path is a dense contraction).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace b200 {

// Programmatic dependent launch (B200_PDL, on by default; measured 2.2 % of a 2^28-key sort): the kernels of one sort are launched with programmatic stream serialisation, so
// the launch latency of a kernel overlaps the tail of its predecessor; every such kernel waits here, before it touches
// anything its predecessors wrote.  A no-op for ordinary launches.
#ifndef B200_PDL
#define B200_PDL 1
#endif
__device__ __forceinline__ void pdl_wait() {
#if B200_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// B200_SEG_CONST=1 (default): the MSD levels recognise a segment whose keys are all equal on the bits still to be sorted and
// finish it by plain copies instead of scattering it at every remaining level (hot keys of duplicate-heavy inputs; cf. CUB's
// short_circuit, lsb/cub/cub/agent/agent_radix_sort_downsweep.cuh:701-724, and the reference's hot-bucket path,
// msb/src/sort/cuda_radix_sort.h:438-447).  Measured on 2^29 Zipf-hashed u64 keys: 29.3 -> 23.7 ms (profiles/r02_cfg4.txt).
#ifndef B200_SEG_CONST
#define B200_SEG_CONST 1
#endif

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 256;

// ---------------------------------------------------------------------------------------------------------
// Order-preserving bit transforms (reference: cub::Traits<T>::TwiddleIn/Out, lsb/cub/cub/util_type.cuh:966-974
// unsigned, :1009-1017 signed, :1079-1089 floating point).  Runtime-parameterised so one kernel instantiation
// per key WIDTH serves u32/i32/f32 (resp. u64/i64/f64); descending order is folded in as a full complement.
// Inside the library keys always travel in the transformed ("twiddled") domain, where plain unsigned
// ascending order is the requested order.
// ---------------------------------------------------------------------------------------------------------
struct Twiddle {
  uint64_t sign_mask;   // top bit for signed-int and float keys, else 0
  uint64_t float_mask;  // all ones for float keys, else 0
  uint64_t flip_mask;   // all ones for descending, else 0
};

template <typename K>
__device__ __forceinline__ K twiddle_in(K k, const Twiddle& t) {
  using S = typename std::make_signed<K>::type;
  const K m = (K)((S)k >> (sizeof(K) * 8 - 1)) & (K)t.float_mask;   // negative float -> complement everything
  return (K)(k ^ (m | (K)t.sign_mask) ^ (K)t.flip_mask);
}
template <typename K>
__device__ __forceinline__ K twiddle_out(K k, const Twiddle& t) {
  using S = typename std::make_signed<K>::type;
  k = (K)(k ^ (K)t.flip_mask);
  const K m = (K)(~(K)((S)k >> (sizeof(K) * 8 - 1))) & (K)t.float_mask;
  return (K)(k ^ (m | (K)t.sign_mask));
}

template <typename K>
__device__ __forceinline__ uint32_t digit_of(K k, int shift, uint32_t mask) {
  return (uint32_t)(k >> shift) & mask;
}

// ---------------------------------------------------------------------------------------------------------
// Decoupled look-back status word (one per tile and digit): [flag:2 | value:30].
// flag 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix.  Flag and value travel in ONE 32-bit word, so
// relaxed loads/stores are enough.  Values are counts relative to the start of the segment's bin, and a launch
// never covers more than 2^30-1 keys per segment bin prefix (the host splits larger inputs into portions).
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t ST_VALUE_MASK = (1u << 30) - 1;
constexpr uint32_t ST_AGG = 1u << 30, ST_PREFIX = 2u << 30;
constexpr uint64_t MAX_PORTION = (1ull << 30) - 1;

__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Warp-level digit matching.  B200SORT_HW_MATCH=1 uses the MATCH.ANY instruction, 0 the ballot loop
// (one VOTE per digit bit).  tools/ubench_rank.cu measures both on sm_100a (results in profiles/).
// ---------------------------------------------------------------------------------------------------------
#ifndef B200SORT_HW_MATCH
#define B200SORT_HW_MATCH 0
#endif
__device__ __forceinline__ unsigned match_digit(unsigned d) {
#if B200SORT_HW_MATCH
  return __match_any_sync(0xffffffffu, d);
#else
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < RADIX_BITS; ++b) {
    const bool p = (d >> b) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, p);
    peers &= p ? m : ~m;
  }
  return peers;
#endif
}

// Exclusive scan of one value per thread over the first 256 threads of the block (8 warps).
// `scratch` = 8 words of shared memory.  Must be called by ALL threads of the block (contains __syncthreads);
// threads >= 256 pass 0 and ignore the result.
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* scratch) {
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  uint32_t inc = v;
  if (w < 8) {                        // warp-uniform: only the eight digit-owner warps take part
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) scratch[w] = inc;
  }
  __syncthreads();
  uint32_t woff = 0;
  if (w < 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t s = scratch[j];
      if ((unsigned)j < w) woff += s;
    }
  }
  __syncthreads();
  return woff + inc - v;
}

// ---------------------------------------------------------------------------------------------------------
// Work descriptors of the MSB levels (device-built; the host never reads them back).
// ---------------------------------------------------------------------------------------------------------
struct Seg {            // a bucket that still needs a counting pass ("non-local bucket" in the reference)
  uint64_t off;         // first key index
  uint32_t cnt;         // number of keys (the MSD engine serves n < 2^32)
  uint32_t flags;       // bit 0: suspected of holding ONE repeated key (all of its parent's keys fell into this one bucket)
};
struct TileDesc {       // one histogram / scatter tile of a segment, self-contained so a tile costs one load
  uint64_t off;         // first key index of the tile
  uint32_t cnt;         // keys in the tile
  uint32_t seg;         // index into the level's Seg list
  uint32_t tile_in_seg;
  uint32_t pad;         // the segment's flags (Seg::flags)
};
struct LocalItem {      // a bucket (or merged run of tiny buckets) that is finished on-chip
  uint64_t off;         // first key index
  uint32_t cnt;         // number of keys (<= the local-sort capacity)
  uint16_t nbits;       // bits [0, nbits) of the twiddled key remain to be sorted
  uint16_t src;         // which of the two ping-pong buffers currently holds the bucket
};

// ---------------------------------------------------------------------------------------------------------
// Destination of a key in the multi-GPU range partition: part = #{ j : splitters[j] <= bucket }, bucket = top `bits` bits.
// A 256-entry table over the top 8 bits of the bucket answers it with one shared-memory load; only the (at most 15) coarse
// cells that contain a splitter in their interior fall back to comparing against the splitters.
// ---------------------------------------------------------------------------------------------------------
struct RangeLut {
  uint32_t split[16];
  uint8_t lut[256];       // low 4 bits: part of the first bucket of the cell; bit 7: a splitter lies inside the cell
  int cshift;             // bucket >> cshift = coarse cell
};
__device__ __forceinline__ void range_lut_build(RangeLut& r, const uint32_t* splitters, int num_parts, int bits) {
  const unsigned tid = threadIdx.x;
  const int cshift = bits > 8 ? bits - 8 : 0;
  if (tid < 256) {
    const uint32_t first = tid << cshift, last = ((tid + 1u) << cshift) - 1u;
    uint32_t d = 0, inside = 0;
    for (int j = 0; j < num_parts - 1; ++j) {
      const uint32_t sp = splitters[j];
      d += sp <= first ? 1u : 0u;
      inside |= (sp > first && sp <= last) ? 1u : 0u;
    }
    r.lut[tid] = (uint8_t)(d | (inside << 7));
  }
  if (tid < 16) r.split[tid] = (int)tid < num_parts - 1 ? splitters[tid] : 0xFFFFFFFFu;
  if (tid == 0) r.cshift = cshift;
}
__device__ __forceinline__ uint32_t range_part(const RangeLut& r, uint32_t bucket, int cshift) {
  const uint32_t e = r.lut[bucket >> cshift];
  if (!(e & 0x80u)) return e;
  uint32_t d = 0;
#pragma unroll
  for (int j = 0; j < 15; ++j) d += bucket >= r.split[j] ? 1u : 0u;
  return d;
}

template <int VB> struct ValType { using type = uint32_t; };
template <> struct ValType<8> { using type = uint64_t; };

}  // namespace b200
