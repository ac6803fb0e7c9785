// msb_sched.cuh -- device-side scheduling of the MSB hybrid sort.  Replaces, without any host round trip,
//   do_fast_merged_compute_subbin_offsets_and_segmented_next_pass_assignments
//       (msb/src/sort/cuda_radix_sort.h:982-1271: prefix sums + classify/merge sub-buckets + emit work lists) and
//   generate_next_pass_block_assignments (HOST code in the reference, msb/src/sort/gpu_radix_sort.cu:29-104).
// Per level: classify_kernel turns every segment's histogram into digit starts, emits the next level's segments
// (sub-buckets larger than the on-chip capacity) and the local-sort work items (everything else; runs of tiny
// neighbours are merged into one item, reference threshold RDXSRT_CFG_MERGE_LOCREC_THRESH,
// cuda_radix_sort_config.h:4); scan_tiles_kernel + fill_descs_kernel cut the next level's segments into tiles.
#pragma once
#include "common.cuh"

namespace b200 {

enum { ERR_SEG_OVERFLOW = 1, ERR_LOCAL_OVERFLOW = 2, ERR_TILE_OVERFLOW = 4 };

struct MsbCounters {            // one small zero-initialised block in the workspace
  uint32_t num_segs[10];        // per level
  uint32_t num_tiles[10];
  uint32_t part_ticket[10];
  uint32_t num_locals[3];       // work lists of the on-chip sorts: ALGO_LSD, ALGO_COUNT, small ALGO_LSD buckets
  uint32_t num_overflow;        // buckets the counting sort handed back
  uint32_t error;
  uint32_t num_bitmap;          // work list of the presence-bitmap sort (large keys-only buckets with <= 16 bits left)
  uint32_t num_direct[2];       // segmented sort: caller segments that fit on chip as they are (large / small on-chip configuration)
  unsigned long long key_or, key_and;     // OR / AND of all transformed keys (level-0 histogram): bits where they agree are constant
  uint32_t probe_single, num_dense;       // (probe_single is read back together with key_or / key_and) the level-0 histogram has one non-empty bucket;
                                          // num_dense: work list of the dense rank sort (large buckets whose keys crowd their cells)
  uint32_t num_overflow_small, pad_;      // handed-back buckets small enough for the 256-thread LSD configuration
};

static __global__ void msb_init_kernel(Seg* segs, MsbCounters* c, uint64_t n) {
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    segs[0].off = 0; segs[0].cnt = (uint32_t)n; segs[0].flags = 0;
    c->num_segs[0] = 1;
    c->key_or = 0ull; c->key_and = ~0ull;
  }
}

// Segmented sort (cub::DeviceSegmentedRadixSort, lsb/cub/cub/device/device_segmented_radix_sort.cuh:140-760): the caller's
// segments ARE the level-0 buckets of the MSD engine.  A segment that fits a CTA's shared memory becomes an on-chip work item
// straight away (its keys are still in caller form: these items get their own launch with the input transform on); a larger
// one enters the level loop like any bucket.  Empty segments (end <= begin) vanish.  Reference semantics: segments do not
// overlap; offsets outside [0, n] raise the error flag and the segment is dropped.
struct SegInitArgs {
  const void* begin; const void* end; uint32_t num_segments; int offset_bytes;
  uint64_t n;
  Seg* segs; uint32_t max_segs;
  LocalItem* direct; LocalItem* direct_small;
  MsbCounters* ctr;
  uint32_t local_cap, small_cap;
  int end_bit;
};
static __global__ void __launch_bounds__(256) seg_init_kernel(const __grid_constant__ SegInitArgs a) {
  pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.ctr->key_or = 0ull; a.ctr->key_and = ~0ull; }
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.num_segments; i += gridDim.x * blockDim.x) {
    long long b, e;
    if (a.offset_bytes == 8) { b = reinterpret_cast<const long long*>(a.begin)[i]; e = reinterpret_cast<const long long*>(a.end)[i]; }
    else { b = reinterpret_cast<const int*>(a.begin)[i]; e = reinterpret_cast<const int*>(a.end)[i]; }
    if (e <= b) continue;
    if (b < 0 || (unsigned long long)e > a.n) { atomicOr(&a.ctr->error, (uint32_t)ERR_SEG_OVERFLOW); continue; }
    const uint64_t cnt = (uint64_t)(e - b);
    if (cnt <= a.local_cap) {
      LocalItem it; it.off = (uint64_t)b; it.cnt = (uint32_t)cnt; it.nbits = (uint16_t)a.end_bit; it.src = 0;
      if (cnt <= a.small_cap) a.direct_small[atomicAdd(&a.ctr->num_direct[1], 1u)] = it;
      else a.direct[atomicAdd(&a.ctr->num_direct[0], 1u)] = it;
    } else {
      const uint32_t slot = atomicAdd(&a.ctr->num_segs[0], 1u);
      if (slot >= a.max_segs) { atomicOr(&a.ctr->error, (uint32_t)ERR_SEG_OVERFLOW); continue; }
      Seg sg; sg.off = (uint64_t)b; sg.cnt = (uint32_t)cnt; sg.flags = 0;
      a.segs[slot] = sg;
    }
  }
}

static __global__ void seg_clamp_kernel(MsbCounters* c, uint32_t max_segs) {
  pdl_wait();
  if (c->num_segs[0] > max_segs) c->num_segs[0] = 0;      // overlapping segments: the error flag is already up, nothing is sorted
}

struct ClassifyArgs {
  const Seg* segs; const uint32_t* num_segs_ptr;
  const uint32_t* seg_hist;     // [segment][256]
  uint64_t* bins;               // [segment][256] out: absolute start of every sub-bucket
  Seg* next_segs; uint32_t* num_next_ptr; uint32_t max_segs;
  LocalItem* locals; uint32_t* num_locals_ptr; uint32_t max_locals;
  LocalItem* locals_small; uint32_t* num_small_ptr; uint32_t small_cap;   // buckets of at most small_cap keys go here (nullptr: none)
  LocalItem* locals_merged; uint32_t* num_merged_ptr;                    // merged runs larger than small_cap go here (nullptr: to `locals`)
  uint32_t* error;
  int shift;                    // bit position of this level's digit; bits [begin_bit, shift) remain below it
  int nb;                       // width of this level's digit (8, or less on the last level of a bit sub-range)
  int last;                     // no bits remain below this digit: every sub-bucket is final after the scatter
  uint32_t local_cap, merge_cap;
  uint32_t merge_small;         // only sub-buckets of at most this many keys are merged with their neighbours (<= merge_cap): two
                                // half-capacity buckets merged cost one more pass over all their keys, apart they cost nothing extra
  uint32_t small_max;           // a bucket that stands alone goes to `locals_small` only below this size (<= small_cap): with more than 8 bits
                                // left the one-shot kernel behind `locals` beats two passes of the small LSD configuration from here on
  int dense_to_merged;          // 1: at most 8 bits remain below this digit -- a bucket that stands alone needs ONE counting pass of
                                // the LSD kernel and goes to `locals_merged`, not to the one-shot list `locals` (its cells would overflow)
  uint32_t out_buf;             // ping-pong buffer the level scatters into
  // B200_SEG_CONST: per-segment OR / AND of the keys (nullptr: off) -- a segment whose keys agree on every bit still to be
  // sorted, [begin_bit, shift + nb), is finished by copy-through items (nbits = begin_bit: zero on-chip passes) in the LSD list
  const unsigned long long* seg_or; const unsigned long long* seg_and;
  LocalItem* locals_copy; uint32_t* num_copy_ptr;
  int begin_bit;
};

constexpr int CLS_WARPS = 4;

static __global__ void __launch_bounds__(CLS_WARPS * 32) classify_kernel(const __grid_constant__ ClassifyArgs a) {
  pdl_wait();
  __shared__ uint32_t s_cnt[CLS_WARPS][RADIX];
  __shared__ uint64_t s_off[CLS_WARPS][RADIX];
  __shared__ LocalItem s_loc[CLS_WARPS][RADIX];
  __shared__ Seg s_seg[CLS_WARPS][RADIX];
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint32_t num_segs = *a.num_segs_ptr;
  for (uint32_t s = blockIdx.x * CLS_WARPS + w; s < num_segs; s += gridDim.x * CLS_WARPS) {
    const Seg sg = a.segs[s];
    // exclusive scan of the 256 counts: lane owns digits 8*lane .. 8*lane+7
    uint32_t c[8]; uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] = a.seg_hist[(uint64_t)s * RADIX + lane * 8 + i]; sum += c[i]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += t;
    }
    uint64_t run = sg.off + (inc - sum);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a.bins[(uint64_t)s * RADIX + lane * 8 + i] = run;
      s_cnt[w][lane * 8 + i] = c[i];
      s_off[w][lane * 8 + i] = run;
      run += c[i];
    }
    __syncwarp();
    if (a.last) continue;             // last digit: every sub-bucket is final after the scatter
#if B200_SEG_CONST
    if (a.seg_or != nullptr && (sg.flags & 1u)) {
      const int top = a.shift + a.nb;
      unsigned long long diff = a.seg_or[s] ^ a.seg_and[s];
      diff &= (top >= 64 ? ~0ull : ((1ull << top) - 1ull)) & ~((1ull << a.begin_bit) - 1ull);
      if (diff == 0ull) {
        // all keys of the segment are equal where it matters: this level's scatter copies it unchanged into out_buf (one digit);
        // nothing below needs sorting -- hand it to the on-chip kernel in capacity-sized chunks that just copy it to the final buffer
        const uint32_t chunks = (uint32_t)(((uint64_t)sg.cnt + a.local_cap - 1) / a.local_cap);
        uint32_t cbase = 0;
        if (lane == 0) cbase = atomicAdd(a.num_copy_ptr, chunks);
        cbase = __shfl_sync(0xffffffffu, cbase, 0);
        if (cbase + chunks > a.max_locals) { if (lane == 0) atomicOr(a.error, (uint32_t)ERR_LOCAL_OVERFLOW); continue; }
        for (uint32_t i = lane; i < chunks; i += 32) {
          LocalItem it; it.off = sg.off + (uint64_t)i * a.local_cap;
          const uint64_t rest = sg.cnt - (uint64_t)i * a.local_cap;
          it.cnt = (uint32_t)(rest < (uint64_t)a.local_cap ? rest : (uint64_t)a.local_cap);
          it.nbits = (uint16_t)a.begin_bit; it.src = (uint16_t)a.out_buf;
          a.locals_copy[cbase + i] = it;
        }
        __syncwarp();
        continue;
      }
    }
#endif
    // Fast path (the common shape of a large uniform sort: 256 sub-buckets that are each too big to share an on-chip item with
    // their neighbour): if no two neighbouring non-empty mergeable sub-buckets fit one item together, the greedy merge below
    // would emit every sub-bucket on its own -- which 32 lanes do in 8 steps instead of one lane in 256.
    {
      bool pair = false;
      uint32_t prev = 0;                                     // count of the previous non-empty digit if it may share an item, else 0
      {
        uint32_t mine = 0; bool has = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c[i]) { has = true; mine = c[i] <= a.merge_small ? c[i] : 0u; }
        const uint32_t hasmask = __ballot_sync(0xffffffffu, has) & ((1u << lane) - 1u);
        const int src = hasmask ? 31 - __clz(hasmask) : 0;
        const uint32_t got = __shfl_sync(0xffffffffu, mine, src);
        prev = hasmask ? got : 0u;                           // (taken from the nearest lower lane that owns a non-empty digit)
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (c[i] == 0) continue;
        const uint32_t cur = c[i] <= a.merge_small ? c[i] : 0u;
        if (prev && cur && prev + cur <= a.merge_cap) pair = true;
        prev = cur;
      }
      if (__ballot_sync(0xffffffffu, pair) == 0u) {
        // single buckets of the one-shot list go to the LSD list instead when one counting pass finishes them (dense_to_merged)
        LocalItem* const single_list = (a.dense_to_merged && a.locals_merged != nullptr) ? a.locals_merged : a.locals;
        uint32_t* const single_cnt = (a.dense_to_merged && a.locals_merged != nullptr) ? a.num_merged_ptr : a.num_locals_ptr;
        uint32_t nl = 0, nm = 0, ng = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (c[i] == 0) continue;
          if (c[i] > a.local_cap) ++ng;
          else if (a.locals_small != nullptr && c[i] <= a.small_max) ++nm;
          else ++nl;
        }
        uint32_t il = nl, im = nm, ig = ng;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t tl = __shfl_up_sync(0xffffffffu, il, o), tm = __shfl_up_sync(0xffffffffu, im, o), tg = __shfl_up_sync(0xffffffffu, ig, o);
          if (lane >= (unsigned)o) { il += tl; im += tm; ig += tg; }
        }
        uint32_t lb = 0, mb = 0, gb = 0;
        if (lane == 31) {
          if (il) lb = atomicAdd(single_cnt, il);
          if (im) mb = atomicAdd(a.num_small_ptr, im);
          if (ig) gb = atomicAdd(a.num_next_ptr, ig);
          if (lb + il > a.max_locals || mb + im > a.max_locals) atomicOr(a.error, (uint32_t)ERR_LOCAL_OVERFLOW);
          if (gb + ig > a.max_segs) atomicOr(a.error, (uint32_t)ERR_SEG_OVERFLOW);
        }
        const uint32_t tl = __shfl_sync(0xffffffffu, il, 31), tm = __shfl_sync(0xffffffffu, im, 31), tg = __shfl_sync(0xffffffffu, ig, 31);
        lb = __shfl_sync(0xffffffffu, lb, 31); mb = __shfl_sync(0xffffffffu, mb, 31); gb = __shfl_sync(0xffffffffu, gb, 31);
        const bool okl = lb + tl <= a.max_locals, okm = mb + tm <= a.max_locals, okg = gb + tg <= a.max_segs;
        uint32_t pl = lb + il - nl, pm = mb + im - nm, pg = gb + ig - ng;
        uint64_t off = sg.off + (inc - sum);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t cd = c[i];
          if (cd != 0) {
            if (cd > a.local_cap) {
              Seg ns; ns.off = off; ns.cnt = cd; ns.flags = cd == sg.cnt ? 1u : 0u;
              if (okg) a.next_segs[pg] = ns;
              ++pg;
            } else {
              LocalItem it; it.off = off; it.cnt = cd; it.nbits = (uint16_t)a.shift; it.src = (uint16_t)a.out_buf;
              if (a.locals_small != nullptr && cd <= a.small_max) { if (okm) a.locals_small[pm] = it; ++pm; }
              else { if (okl) single_list[pl] = it; ++pl; }
            }
          }
          off += cd;
        }
        __syncwarp();
        continue;
      }
    }
    // classify + merge, serial over the 256 digits (lane 0), staged in shared memory
    uint32_t nloc = 0, nseg = 0, nsml = 0, nmrg = 0;     // big items fill s_loc[w] from the front, small ones from the back;
    LocalItem* s_mrg = reinterpret_cast<LocalItem*>(&s_seg[w][0]);   // merged runs share s_seg[w] with the segments, from the back
    static_assert(sizeof(LocalItem) == sizeof(Seg), "merged items are staged in the segment array");
    if (lane == 0) {
      uint64_t pend_off = 0; uint32_t pend_sum = 0, pend_n = 0;
      auto flush = [&]() {
        if (pend_n) {
          LocalItem it; it.off = pend_off; it.cnt = pend_sum;
          it.nbits = (uint16_t)(pend_n > 1 ? a.shift + a.nb : a.shift); it.src = (uint16_t)a.out_buf;
          if (a.locals_small != nullptr && it.cnt <= (pend_n > 1 ? a.small_cap : a.small_max)) s_loc[w][RADIX - 1 - nsml++] = it;
          else if (a.locals_merged != nullptr && (pend_n > 1 || a.dense_to_merged)) s_mrg[RADIX - 1 - nmrg++] = it;
          else s_loc[w][nloc++] = it;
          pend_n = 0; pend_sum = 0;
        }
      };
      for (int d = 0; d < RADIX; ++d) {
        const uint32_t cd = s_cnt[w][d];
        if (cd == 0) continue;
        if (cd > a.local_cap) {
          flush();
          Seg ns; ns.off = s_off[w][d]; ns.cnt = cd; ns.flags = cd == sg.cnt ? 1u : 0u;      // everything fell into one bucket: one repeated key?
          s_seg[w][nseg++] = ns;
        } else if (cd > a.merge_small) {     // stands alone
          flush();
          LocalItem it; it.off = s_off[w][d]; it.cnt = cd; it.nbits = (uint16_t)a.shift; it.src = (uint16_t)a.out_buf;
          if (a.locals_small != nullptr && cd <= a.small_max) s_loc[w][RADIX - 1 - nsml++] = it;
          else if (a.locals_merged != nullptr && a.dense_to_merged) s_mrg[RADIX - 1 - nmrg++] = it;
          else s_loc[w][nloc++] = it;
        } else {
          if (pend_n && pend_sum + cd > a.merge_cap) flush();
          if (pend_n == 0) pend_off = s_off[w][d];
          pend_sum += cd; ++pend_n;
        }
      }
      flush();
    }
    nloc = __shfl_sync(0xffffffffu, nloc, 0);
    nsml = __shfl_sync(0xffffffffu, nsml, 0);
    nmrg = __shfl_sync(0xffffffffu, nmrg, 0);
    nseg = __shfl_sync(0xffffffffu, nseg, 0);
    __syncwarp();
    uint32_t lbase = 0, sbase = 0, mbase = 0, gbase = 0;
    if (lane == 0) {
      if (nloc) lbase = atomicAdd(a.num_locals_ptr, nloc);
      if (nsml) mbase = atomicAdd(a.num_small_ptr, nsml);
      if (nmrg) gbase = atomicAdd(a.num_merged_ptr, nmrg);
      if (nseg) sbase = atomicAdd(a.num_next_ptr, nseg);
    }
    lbase = __shfl_sync(0xffffffffu, lbase, 0);
    mbase = __shfl_sync(0xffffffffu, mbase, 0);
    gbase = __shfl_sync(0xffffffffu, gbase, 0);
    sbase = __shfl_sync(0xffffffffu, sbase, 0);
    if (lbase + nloc > a.max_locals) { if (lane == 0) atomicOr(a.error, (uint32_t)ERR_LOCAL_OVERFLOW); nloc = 0; }
    if (mbase + nsml > a.max_locals) { if (lane == 0) atomicOr(a.error, (uint32_t)ERR_LOCAL_OVERFLOW); nsml = 0; }
    if (gbase + nmrg > a.max_locals) { if (lane == 0) atomicOr(a.error, (uint32_t)ERR_LOCAL_OVERFLOW); nmrg = 0; }
    if (sbase + nseg > a.max_segs) { if (lane == 0) atomicOr(a.error, (uint32_t)ERR_SEG_OVERFLOW); nseg = 0; }
    for (uint32_t i = lane; i < nloc; i += 32) a.locals[lbase + i] = s_loc[w][i];
    for (uint32_t i = lane; i < nsml; i += 32) a.locals_small[mbase + i] = s_loc[w][RADIX - 1 - i];
    for (uint32_t i = lane; i < nmrg; i += 32) a.locals_merged[gbase + i] = s_mrg[RADIX - 1 - i];
    for (uint32_t i = lane; i < nseg; i += 32) a.next_segs[sbase + i] = s_seg[w][i];
    __syncwarp();
  }
}

// Exclusive scan of tiles-per-segment -> tile_base[0..num_segs], total -> *num_tiles_ptr.  ONE CTA.
constexpr int SCAN_THREADS = 1024;
static __global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const Seg* segs, const uint32_t* num_segs_ptr, uint32_t* tile_base,
                                                                   uint32_t* num_tiles_ptr, uint32_t max_tiles, uint32_t* error, int tile) {
  pdl_wait();
  __shared__ uint32_t s_w[32];
  const uint32_t ns = min(*num_segs_ptr, 0x7fffffffu);
  const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const uint32_t per = (ns + SCAN_THREADS - 1) / SCAN_THREADS;
  const uint32_t s0 = min(per * tid, ns), s1 = min(s0 + per, ns);
  uint32_t sum = 0;
  for (uint32_t s = s0; s < s1; ++s) sum += (uint32_t)(((uint64_t)segs[s].cnt + tile - 1) / tile);
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (unsigned)o) inc += t;
  }
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t v = s_w[lane], vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
      if (lane >= (unsigned)o) vi += t;
    }
    s_w[lane] = vi - v;
    if (lane == 31) {
      uint32_t total = vi;
      if (total > max_tiles) { atomicOr(error, (uint32_t)ERR_TILE_OVERFLOW); total = 0; }
      *num_tiles_ptr = total;
      tile_base[ns] = total;
    }
  }
  __syncthreads();
  uint32_t run = s_w[w] + inc - sum;
  for (uint32_t s = s0; s < s1; ++s) {
    tile_base[s] = run;
    run += (uint32_t)(((uint64_t)segs[s].cnt + tile - 1) / tile);
  }
}

static __global__ void fill_descs_kernel(const Seg* segs, const uint32_t* tile_base, const uint32_t* num_segs_ptr, const uint32_t* num_tiles_ptr,
                                         TileDesc* descs, int tile) {
  pdl_wait();
  const uint32_t ns = *num_segs_ptr, nt = *num_tiles_ptr;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    uint32_t lo = 0, hi = ns;           // largest s with tile_base[s] <= t
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (tile_base[mid] <= t) lo = mid; else hi = mid;
    }
    const Seg sg = segs[lo];
    TileDesc td; td.seg = lo; td.tile_in_seg = t - tile_base[lo]; td.pad = sg.flags;
    const uint64_t rel = (uint64_t)td.tile_in_seg * tile;
    td.off = sg.off + rel;
    td.cnt = (uint32_t)(sg.cnt - rel < (uint64_t)tile ? sg.cnt - rel : (uint64_t)tile);
    descs[t] = td;
  }
}

}  // namespace b200
