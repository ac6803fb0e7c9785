// sort_impl.cuh -- host orchestration (stream-ordered, no host<->device synchronisation, no allocation when the
// caller supplies the temporary storage).
//   lsb_sort_impl : stable LSD sort  = cub::DeviceRadixSort::{SortKeys,SortPairs}[Descending] call shape
//                   (lsb/cub/cub/device/device_radix_sort.cuh:147-781; DispatchRadixSort::InvokePasses,
//                   lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:1050-1159), built as: one histogram read for all
//                   digits -> one onesweep-style partition launch per digit.
//   msb_sort_impl : MSB hybrid sort = rdxsrt_unstable_sort (msb/src/sort/gpu_radix_sort.h:187-507), built as:
//                   per level [segment histograms -> classify -> partition -> next tile list], then ONE local-sort
//                   launch over every bucket that fits on chip.  All scheduling stays on the device.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <utility>
#include "hist.cuh"
#include "local_sort.cuh"
#include "local_bitmap.cuh"
#include "local_rank.cuh"
#include "msb_sched.cuh"
#include "scatter.cuh"
#include "sort_api.h"

namespace b200 {

#define B200_CHECK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return e__; } while (0)

// Everything cached per process is cached PER DEVICE (a process may drive several GPUs, like callers of the CUB and reference
// entry points can): SM counts, the persistent-grid sizes (and the cudaFuncSetAttribute call that goes with them), probe events.
constexpr int MAX_DEVICES = 64;
inline int current_device() { int dev = 0; cudaGetDevice(&dev); return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0; }
inline int num_sms() {
  static int sms[MAX_DEVICES] = {};
  const int dev = current_device();
  if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev];
}

// One launch of the sort's kernel chain: ordinary, or (B200_PDL) with programmatic stream serialisation.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
#if B200_PDL
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  note_launch();
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
#else
  kernel<<<grid, block, smem, s>>>(std::forward<Args>(args)...); note_launch();
  return cudaSuccess;
#endif
}

template <typename K, int VB>
struct Cfg {
  // scatter tile: 512 threads x IPT keys, sized so that two CTAs (double-buffered key + value staging, rank arrays) share an SM
#ifndef B200_SCATTER_THREADS
#define B200_SCATTER_THREADS 512
#endif
#ifndef B200_IPT_NUM
#define B200_IPT_NUM 1
#define B200_IPT_DEN 1
#endif
#ifndef B200_HIST_TICKET
#define B200_HIST_TICKET 1
#endif
#ifndef B200_LOCAL_IPT32
#define B200_LOCAL_IPT32 12
#endif
#ifndef B200_LOCAL_THREADS32
#define B200_LOCAL_THREADS32 384
#endif
  static constexpr int THREADS = B200_SCATTER_THREADS;
  static constexpr int IPT = (sizeof(K) == 4 ? (VB == 0 ? 16 : VB == 4 ? 8 : 5) : (VB == 0 ? 8 : VB == 4 ? 5 : 4)) * B200_IPT_NUM / B200_IPT_DEN;
#ifndef B200_SCATTER_OCC
#define B200_SCATTER_OCC (1024 / B200_SCATTER_THREADS)
#endif
  static constexpr int OCC = B200_SCATTER_OCC;        // scatter CTAs per SM the launch bounds ask for
  static constexpr int TILE = THREADS * IPT;
  // local sort: capacity = the largest bucket that is finished on chip (everything larger gets another level)
  static constexpr int LOCAL_THREADS = sizeof(K) == 4 ? B200_LOCAL_THREADS32 : (VB == 0 ? 768 : 512);
  static constexpr int LOCAL_IPT = sizeof(K) == 4 ? (VB == 8 ? 8 : B200_LOCAL_IPT32) : (VB == 0 ? 12 : 8);
  static constexpr int LOCAL_CAP = LOCAL_THREADS * LOCAL_IPT;
  // runs of small neighbouring buckets are merged into one on-chip item (which then also sorts the current digit): up to the
  // full capacity where the LSD kernel finishes them (one more 8-bit pass beats many half-empty items), up to a quarter of it
  // where the one-shot counting sort does (its cells are indexed by the leading bits, which a merged run barely varies)
  static constexpr uint32_t MERGE_CAP = LOCAL_CAP / 4;
  // second on-chip configuration for small buckets (merged tiny buckets, skewed inputs): fewer threads, four CTAs per SM
  static constexpr int SMALL_THREADS = 256;
  static constexpr int SMALL_IPT = 6;
  static constexpr int SMALL_CAP = SMALL_THREADS * SMALL_IPT;
  // handed-back buckets (skewed digits, several LSD passes): the 256-thread configuration takes twice as many keys per thread, so
  // that most of them (merged runs of up to a quarter of the large capacity, single buckets of a few thousand keys) run three
  // to an SM instead of alone in the large configuration (measured per key and pass: 1.5x; AND-entropy u64 keys 36.6 -> 27.1 ms)
  static constexpr int SKEW_SMALL_IPT = (sizeof(K) + VB <= 12) ? 12 : SMALL_IPT;
  static constexpr int SKEW_SMALL_CAP = SMALL_THREADS * SKEW_SMALL_IPT;
};

// Persistent-grid size of a kernel: resident CTAs per SM x SMs (queried once per instantiation).
template <typename KernelT>
inline cudaError_t persistent_grid(KernelT kernel, int threads, size_t smem, int* grid) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  *grid = occ * num_sms();
  return cudaSuccess;
}

template <typename K, int VB>
inline cudaError_t launch_scatter_fast(const ScatterArgs& a, uint32_t tiles_hint, cudaStream_t s) {
  using C = Cfg<K, VB>;
  auto kernel = scatter_fast_kernel<K, VB, C::THREADS, C::IPT, C::OCC>;
  constexpr size_t smem = sizeof(FastSmem<K, VB, C::THREADS, C::IPT>);
  static int grids[MAX_DEVICES] = {};
  int& grid = grids[current_device()];
  if (!grid) B200_CHECK(persistent_grid(kernel, C::THREADS, smem, &grid));
  const int g = (int)std::min<uint64_t>((uint64_t)grid, std::max<uint32_t>(tiles_hint, 1u));
  ProfScope prof("scatter", s);
  launch_k(kernel, g, C::THREADS, smem, s, a);
  return cudaGetLastError();
}

template <typename K, int VB>
inline cudaError_t launch_scatter_stable_fast(const ScatterArgs& a, uint32_t tiles_hint, cudaStream_t s) {
  using C = Cfg<K, VB>;
  auto kernel = scatter_stable_fast_kernel<K, VB, C::THREADS, C::IPT, C::OCC>;
  constexpr size_t smem = sizeof(StableFastSmem<K, VB, C::THREADS, C::IPT>);
  static_assert(smem <= 113 * 1024, "two stable scatter CTAs must fit one SM");
  static int grids[MAX_DEVICES] = {};
  int& grid = grids[current_device()];
  if (!grid) B200_CHECK(persistent_grid(kernel, C::THREADS, smem, &grid));
  const int g = (int)std::min<uint64_t>((uint64_t)grid, std::max<uint32_t>(tiles_hint, 1u));
  ProfScope prof("scatter_stable", s);
  launch_k(kernel, g, C::THREADS, smem, s, a);
  return cudaGetLastError();
}

template <typename K, int VB>
inline cudaError_t launch_scatter_exchange(const ScatterArgs& a, uint32_t tiles_hint, cudaStream_t s) {
  using C = Cfg<K, VB>;
  auto kernel = scatter_stable_fast_kernel<K, VB, C::THREADS, C::IPT, C::OCC, true>;
  constexpr size_t smem = sizeof(StableFastSmem<K, VB, C::THREADS, C::IPT, true>);
  static_assert(smem <= 113 * 1024, "two exchange scatter CTAs must fit one SM");
  static int grids[MAX_DEVICES] = {};
  int& grid = grids[current_device()];
  if (!grid) B200_CHECK(persistent_grid(kernel, C::THREADS, smem, &grid));
  const int g = (int)std::min<uint64_t>((uint64_t)grid, std::max<uint32_t>(tiles_hint, 1u));
  ProfScope prof("exchange_scatter", s);
  launch_k(kernel, g, C::THREADS, smem, s, a);
  return cudaGetLastError();
}

template <typename K, int VB, int MODE, bool ORD>
inline cudaError_t launch_scatter(const ScatterArgs& a, uint32_t tiles_hint, cudaStream_t s) {
  using C = Cfg<K, VB>;
  if (MODE == MODE_SEG && !ORD && C::THREADS >= 2 * RADIX) return launch_scatter_fast<K, VB>(a, tiles_hint, s);
  if (MODE == MODE_SEG && ORD && C::THREADS >= 2 * RADIX) return launch_scatter_stable_fast<K, VB>(a, tiles_hint, s);
  auto kernel = scatter_kernel<K, VB, C::THREADS, C::IPT, C::OCC, MODE, ORD>;
  constexpr size_t smem = sizeof(ScatterSmem<K, VB, C::THREADS, C::IPT, MODE, ORD>);
  static_assert(smem <= (C::OCC >= 2 ? 113 : 227) * 1024, "the scatter CTAs of one SM must fit its 228 KB of shared memory");
  static int grids[MAX_DEVICES] = {};
  int& grid = grids[current_device()];
  if (!grid) B200_CHECK(persistent_grid(kernel, C::THREADS, smem, &grid));
  const int g = (int)std::min<uint64_t>((uint64_t)grid, std::max<uint32_t>(tiles_hint, 1u));
  ProfScope prof(MODE == MODE_RANGE ? "range_partition" : (MODE == MODE_LSB ? "scatter_onesweep" : (ORD ? "scatter_stable" : "scatter")), s);
  kernel<<<g, C::THREADS, smem, s>>>(a); note_launch();
  return cudaGetLastError();
}

template <typename K, int VB, int ALGO, bool STABLE, bool SMALL = false, bool SKEW = false>
inline cudaError_t launch_local(const LocalArgs& a, uint32_t items_hint, cudaStream_t s) {
  using C = Cfg<K, VB>;
  constexpr int THREADS = SMALL ? C::SMALL_THREADS : C::LOCAL_THREADS;
  constexpr int IPT = SMALL ? (SKEW ? C::SKEW_SMALL_IPT : C::SMALL_IPT) : C::LOCAL_IPT;
  auto kernel = local_sort_kernel<K, VB, THREADS, IPT, ALGO, STABLE, SKEW>;
  constexpr size_t smem = sizeof(LocalSmem<K, VB, THREADS, IPT, ALGO>);
  static_assert(smem <= 227 * 1024, "local sort exceeds the 227 KB shared-memory limit");
  static int grids[MAX_DEVICES] = {};
  int& grid = grids[current_device()];
  if (!grid) B200_CHECK(persistent_grid(kernel, THREADS, smem, &grid));
  const int g = (int)std::min<uint64_t>((uint64_t)grid, std::max<uint32_t>(items_hint, 1u));
  ProfScope prof(ALGO == ALGO_LSD ? (SMALL ? "local_sort_lsd_small" : "local_sort_lsd") : "local_sort_count", s);
  launch_k(kernel, g, THREADS, smem, s, a);
  return cudaGetLastError();
}

// Keys-only buckets with <= 16 undecided bits where the sort covers the whole key: presence-bitmap sort (local_bitmap.cuh).
#ifndef B200_BITMAP_THREADS
#define B200_BITMAP_THREADS 256
#endif
#ifndef B200_BITMAP_OCC
#define B200_BITMAP_OCC 4
#endif
template <typename K, int VB>
inline cudaError_t launch_bitmap(const LocalArgs& a, uint32_t items_hint, cudaStream_t s) {
  using C = Cfg<K, VB>;
  auto kernel = bitmap_sort_kernel<K, B200_BITMAP_THREADS, C::LOCAL_CAP, B200_BITMAP_OCC>;
  constexpr size_t smem = sizeof(BitmapSmem<K, B200_BITMAP_THREADS, C::LOCAL_CAP>);
  static int grids[MAX_DEVICES] = {};
  int& grid = grids[current_device()];
  if (!grid) B200_CHECK(persistent_grid(kernel, B200_BITMAP_THREADS, smem, &grid));
  const int g = (int)std::min<uint64_t>((uint64_t)grid, std::max<uint32_t>(items_hint, 1u));
  ProfScope prof("local_sort_bitmap", s);
  launch_k(kernel, g, B200_BITMAP_THREADS, smem, s, a);
  return cudaGetLastError();
}

// 4-byte keys, buckets with <= 16 undecided bits: one-shot rank by presence bits (local_rank.cuh).
#ifndef B200_RANK_THREADS
#define B200_RANK_THREADS 512      // pairs
#endif
#ifndef B200_RANK_THREADS0
#define B200_RANK_THREADS0 256     // keys-only
#endif
#ifndef B200_RANK_OCC
#define B200_RANK_OCC 0            // 0: as many CTAs per SM as shared memory and the thread limit allow
#endif
#ifndef B200_RANK_OCC0
#define B200_RANK_OCC0 0
#endif
#ifndef B200_RANK_MIN
#define B200_RANK_MIN 1537        // smallest single bucket (9-16 bits left) routed to the rank kernel instead of the small LSD configuration
#endif
template <typename K, int VB, bool STABLE, bool DENSE = false>
inline cudaError_t launch_rank(const LocalArgs& a, uint32_t items_hint, cudaStream_t s) {
  using C = Cfg<K, VB>;
  constexpr int THREADS = (VB || DENSE) ? B200_RANK_THREADS : B200_RANK_THREADS0, IPT = C::LOCAL_CAP / THREADS;
  static_assert(THREADS * IPT == C::LOCAL_CAP, "the rank kernel's capacity is the on-chip capacity of the configuration");
  constexpr size_t smem = sizeof(RankSmem<K, VB, THREADS, IPT, STABLE>);
  constexpr int OCC_SET = (VB || DENSE) ? B200_RANK_OCC : B200_RANK_OCC0;
  constexpr int OCC = OCC_SET ? OCC_SET : (int)std::min<size_t>((227 * 1024) / (smem + 1024), THREADS >= 512 ? 2 : 2048 / THREADS);      // (512 threads: 64 registers each)
  auto kernel = rank_sort_kernel<K, VB, THREADS, IPT, OCC, STABLE, DENSE>;
  static int grids[MAX_DEVICES] = {};
  int& grid = grids[current_device()];
  if (!grid) B200_CHECK(persistent_grid(kernel, THREADS, smem, &grid));
  const int g = (int)std::min<uint64_t>((uint64_t)grid, std::max<uint32_t>(items_hint, 1u));
  ProfScope prof(DENSE ? "local_sort_rank_dense" : "local_sort_rank", s);
  launch_k(kernel, g, THREADS, smem, s, a);
  return cudaGetLastError();
}

// Tiny helper kernels --------------------------------------------------------------------------------------------
static __global__ void single_item_kernel(LocalItem* item, uint32_t* num_items, uint32_t cnt, int nbits) {
  pdl_wait();
  LocalItem it; it.off = 0; it.cnt = cnt; it.nbits = (uint16_t)nbits; it.src = 0;
  *item = it; *num_items = 1;
}
// Zeroes the rows of the per-level arrays that the level will actually use.
static __global__ void level_prep_kernel(uint32_t* seg_hist, const uint32_t* num_segs_ptr, unsigned long long* seg_or, unsigned long long* seg_and) {
  pdl_wait();
  const uint64_t nh = (uint64_t)*num_segs_ptr * RADIX / 4;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (uint64_t i = t; i < nh; i += stride) reinterpret_cast<uint4*>(seg_hist)[i] = z;
  if (seg_or != nullptr)
    for (uint64_t i = t; i < (uint64_t)*num_segs_ptr; i += stride) { seg_or[i] = 0ull; seg_and[i] = ~0ull; }
}

struct Carver {      // sub-allocates the caller's temporary storage, 256-byte aligned
  char* base; size_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<char*>(p)) {}
  template <typename T> T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
  size_t total() const { return (off + 255) & ~(size_t)255; }
};

// ===============================================================================================================
// The MSD engine shared by both entry points: per level [tile histograms -> group carry -> classify -> scatter ->
// next level's tile list], then ONE on-chip sort launch over every bucket that fits a CTA's shared memory.
//   ORDERED = false : unstable (MSB hybrid sort, rdxsrt_unstable_sort): atomic in-tile ranking, one-shot counting local sort
//   ORDERED = true  : stable (the cub::DeviceRadixSort contract): match-based in-tile ranking, stable LSD local sort.
// Tiles never wait for each other: a tile's destination = sub-bucket start + counts of earlier tiles, all known before
// the scatter starts, so placement is deterministic and (with stable ranking) the whole sort is stable.
// bufk/bufv: up to three ping-pong buffers; buffer 0 holds the input; level L scatters into out_of_level(L); buckets are
// finished into buffer `fin`.  Sorts on bits [begin_bit, end_bit) of the transformed key.  n < 2^32.
// ===============================================================================================================
struct MsdWorkspace {
  MsbCounters* ctr; Seg* segs0; Seg* segs1; uint32_t* tile_base; TileDesc* descs; uint32_t* seg_hist; uint64_t* bins;
  uint32_t* tile_off; uint16_t* tile_cnt; uint32_t* group_tail; uint32_t* group_flag; uint32_t* carry; LocalItem* locals[7];   // LSD list, counting list, overflow, small-bucket LSD list, rank list, dense rank list, small overflow
  uint32_t max_segs, max_tiles, max_locals, max_groups;
  unsigned long long* seg_or; unsigned long long* seg_and;      // B200_SEG_CONST: per-segment OR / AND of the keys
};

template <typename K, int VB>
inline void msd_carve(Carver& cv, uint64_t n, MsdWorkspace& w) {
  using C = Cfg<K, VB>;
  constexpr int LEVELS = sizeof(K);
  w.max_segs = (uint32_t)(n / C::LOCAL_CAP) + 2;
  w.max_tiles = (uint32_t)(n / C::TILE) + w.max_segs + 1;
  w.max_groups = w.max_tiles / HIST_GROUP + 1;
  w.max_locals = (uint32_t)std::min<uint64_t>(4 * n / C::MERGE_CAP + 4ull * LEVELS * w.max_segs + 16, 0x7fffffffu);
  w.ctr = cv.take<MsbCounters>(1);
  w.segs0 = cv.take<Seg>(w.max_segs);
  w.segs1 = cv.take<Seg>(w.max_segs);
  w.tile_base = cv.take<uint32_t>(w.max_segs + 1);
  w.descs = cv.take<TileDesc>(w.max_tiles);
  w.seg_hist = cv.take<uint32_t>((size_t)w.max_segs * RADIX);
  w.bins = cv.take<uint64_t>((size_t)w.max_segs * RADIX);
  w.tile_off = cv.take<uint32_t>((size_t)w.max_tiles * RADIX);
  w.tile_cnt = cv.take<uint16_t>((size_t)w.max_tiles * RADIX);
  w.group_tail = cv.take<uint32_t>((size_t)w.max_groups * RADIX);
  w.group_flag = cv.take<uint32_t>(w.max_groups);
  w.carry = cv.take<uint32_t>((size_t)w.max_groups * RADIX);
  for (int i = 0; i < 7; ++i) w.locals[i] = cv.take<LocalItem>(w.max_locals);
#if B200_SEG_CONST
  w.seg_or = cv.take<unsigned long long>(w.max_segs);
  w.seg_and = cv.take<unsigned long long>(w.max_segs);
#endif
}

// Pinned landing zone for the key-range probe: key_or, key_and, {probe_single, -} (one per host thread).
inline unsigned long long* probe_buffer() {
  static thread_local unsigned long long* p = nullptr;
  if (!p && cudaHostAlloc(reinterpret_cast<void**>(&p), 4 * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess) p = nullptr;
  return p;
}
inline cudaEvent_t probe_event() {
  static thread_local cudaEvent_t evs[MAX_DEVICES] = {};
  cudaEvent_t& ev = evs[current_device()];
  if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) ev = nullptr;
  return ev;
}
// Runs right after the level-0 histogram of a probing sort.  If the keys agree on leading bits of the sorted window, the sort
// that is already enqueued behind this kernel would waste whole levels on single-bucket digits: empty its work lists so that
// every kernel behind finds nothing to do (they all read their counts from device memory); the host, which learns the same
// from the read-back, then enqueues the sort again on the narrower window.
//   FROM_HIST: the OR / AND of the leading digit follow from WHICH level-0 buckets are non-empty (no per-key work in the
//   histogram kernel); the bits below the digit count as differing.  A single non-empty bucket says nothing about the lower
//   bits: key_or = key_and = 0 and probe_single = 1 ask the host for a round with the exact per-key OR / AND.
template <bool FROM_HIST>
static __global__ void __launch_bounds__(RADIX) probe_eval_kernel(MsbCounters* c, const uint32_t* seg_hist, int shift0, int begin_bit, int end_bit) {
  pdl_wait();
  __shared__ uint32_t s_or, s_and, s_cnt;
  if (FROM_HIST) {
    if (threadIdx.x == 0) { s_or = 0; s_and = 0xFFFFFFFFu; s_cnt = 0; }
    __syncthreads();
    if (seg_hist[threadIdx.x] != 0) { atomicOr(&s_or, threadIdx.x); atomicAnd(&s_and, threadIdx.x); atomicAdd(&s_cnt, 1u); }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  if (FROM_HIST) {
    const unsigned long long low = (1ull << shift0) - 1ull;
    c->probe_single = s_cnt <= 1 ? 1u : 0u;
    c->key_or = s_cnt <= 1 ? 0ull : (((unsigned long long)s_or << shift0) | low);
    c->key_and = s_cnt <= 1 ? 0ull : ((unsigned long long)s_and << shift0);
  }
  unsigned long long diff = c->key_or ^ c->key_and;
  diff &= (end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1ull)) & ~((1ull << begin_bit) - 1ull);
  if (end_bit > begin_bit && ((diff >> (end_bit - 1)) & 1ull) == 0ull) { c->num_tiles[0] = 0; c->num_segs[0] = 0; }
}
constexpr uint64_t PROBE_MIN_ITEMS = 1ull << 22;     // below this the read-back and the host's event wait cost more than the probe can save

// Caller-defined segments (segmented sort): offsets on the device, plus two work lists for the segments that fit on chip as they are.
struct SegInput {
  const void* begin; const void* end; uint32_t num_segments; int offset_bytes;
  LocalItem* direct; LocalItem* direct_small;
};

// fin_in: buffer the result must land in (-1: the engine picks bufk[levels & 1] of the two ping-pong buffers); *fin_out says where it is.
// segin != nullptr: every caller segment is sorted on its own (the segments replace the single level-0 bucket [0, n)).
template <typename K, int VB, bool ORDERED>
cudaError_t msd_sort_run(const MsdWorkspace& w, void* const bufk[3], void* const bufv[3], int nbuf, int fin_in, int* fin_out, uint64_t n,
                         const Twiddle& tw, int begin_bit, int end_bit, cudaStream_t s, const SegInput* segin = nullptr) {
  using C = Cfg<K, VB>;
  using V = typename ValType<VB>::type;
  const int sms = num_sms();
  MsbCounters* ctr = w.ctr;
  const int twid = (tw.sign_mask | tw.float_mask | tw.flip_mask) != 0 ? 1 : 0;     // unsigned ascending keys: the transform is the identity
  int levels = (end_bit - begin_bit + 7) / 8;
  int fin = fin_in >= 0 ? fin_in : (levels & 1);
  // level L scatters from in_buf(L) to out_buf(L); the LAST possible level must land in `fin`
  auto out_buf = [&](int L) -> int {
    if (nbuf == 2) return fin_in >= 0 ? (((levels - 1 - L) & 1) == 0 ? fin : fin ^ 1) : (L + 1) & 1;
    return ((levels - 1 - L) & 1) == 0 ? fin : (fin == 1 ? 2 : 1);      // never the input buffer 0
  };
  auto in_buf = [&](int L) -> int { return L == 0 ? 0 : out_buf(L - 1); };

  LocalArgs la{};
  for (int i = 0; i < 3; ++i) { la.keys[i] = bufk[i < nbuf ? i : 0]; la.vals[i] = bufv[i < nbuf ? i : 0]; }
  la.items = w.locals[0]; la.num_items_ptr = &ctr->num_locals[0];
  la.overflow = w.locals[2]; la.num_overflow_ptr = &ctr->num_overflow;
  la.overflow_small = w.locals[6]; la.num_overflow_small_ptr = &ctr->num_overflow_small; la.overflow_small_cap = C::SKEW_SMALL_CAP;
  la.max_items = w.max_locals; la.error_ptr = &ctr->error;
  // the sort covers the whole key: buckets whose undecided bits are equal hold equal keys, so the on-chip sort may rebuild keys from cells
  const bool whole_key = begin_bit == 0 && end_bit == (int)sizeof(K) * 8;
  // which kernel finishes the large buckets with <= 16 bits left (measurement switch B200SORT_LOCAL=rank|bitmap|old)
  bool use_rank = sizeof(K) == 4;
  bool use_bitmap = false;
  { static const char* e = getenv("B200SORT_LOCAL");
    if (e && e[0] == 'b') { use_rank = false; use_bitmap = !ORDERED && VB == 0 && sizeof(K) == 4 && whole_key; }
    if (e && e[0] == 'o') use_rank = false; }
  la.tw_out = twid; la.begin_bit = begin_bit; la.tw = tw;

  if (segin == nullptr && n <= (uint64_t)C::LOCAL_CAP) {     // fits one CTA: a single on-chip sort straight into the final buffer
    if (fin_in < 0) fin = 0;
    if (fin_out) *fin_out = fin;
    la.keys_final = bufk[fin]; la.vals_final = bufv[fin];
    launch_k(single_item_kernel, 1, 1, 0, s, w.locals[0], &ctr->num_locals[0], (uint32_t)n, end_bit);
    la.tw_in = twid;
    return launch_local<K, VB, ALGO_LSD, ORDERED>(la, 1, s);
  }

  // Key-range probe: the level-0 histogram also ORs / ANDs every key.  Leading bits on which all keys agree cannot influence
  // the order, so the digit windows start below them (small-range keys, e.g. indices below 2^20 in 64-bit keys, would
  // otherwise spend whole sweeps on single-bucket levels).  The sort is enqueued optimistically on the full window; a one-thread
  // kernel behind the level-0 histogram empties the work lists if the window turns out too wide, and the host -- after one
  // event wait at the END of the enqueue, when the device is busy -- enqueues the narrower sort.  Skipped for small inputs and
  // while the stream is being captured into a CUDA graph.
  // probe = 1: OR / AND of the leading digit from the level-0 histogram (free); 2: exact per-key OR / AND; 0: none
  int probe = (n >= PROBE_MIN_ITEMS && levels > 1 && segin == nullptr) ? 1 : 0;
  if (!g_key_range_probe) probe = 0;                                                          // b200_set_key_range_probe(0): strictly host-asynchronous calls
  { static const char* e = getenv("B200SORT_PROBE"); if (e && e[0] == '0') probe = 0; }      // measurement switch
  if (probe) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone || probe_buffer() == nullptr || probe_event() == nullptr) probe = 0;
  }
  for (;;) {      // the optimistic enqueue; then, only if the probe asks: (an exact-probe round and) the narrowed sort
    B200_CHECK(cudaMemsetAsync(ctr, 0, sizeof(MsbCounters), s));
    {
      ProfScope prof("msb_sched", s);
      if (segin != nullptr) {
        SegInitArgs sa{};
        sa.begin = segin->begin; sa.end = segin->end; sa.num_segments = segin->num_segments; sa.offset_bytes = segin->offset_bytes;
        sa.n = n; sa.segs = w.segs0; sa.max_segs = w.max_segs; sa.direct = segin->direct; sa.direct_small = segin->direct_small;
        sa.ctr = ctr; sa.local_cap = C::LOCAL_CAP; sa.small_cap = C::SMALL_CAP; sa.end_bit = end_bit;
        const int sgrid = (int)std::min<uint32_t>((segin->num_segments + 255) / 256, (uint32_t)sms * 8);
        launch_k(seg_init_kernel, std::max(sgrid, 1), 256, 0, s, sa);
        launch_k(seg_clamp_kernel, 1, 1, 0, s, ctr, w.max_segs);
      } else {
        launch_k(msb_init_kernel, 1, 32, 0, s, w.segs0, ctr, n);
      }
      launch_k(scan_tiles_kernel, 1, SCAN_THREADS, 0, s, w.segs0, &ctr->num_segs[0], w.tile_base, &ctr->num_tiles[0], w.max_tiles, &ctr->error, C::TILE);
      launch_k(fill_descs_kernel, sms * 2, 256, 0, s, w.segs0, w.tile_base, &ctr->num_segs[0], &ctr->num_tiles[0], w.descs, C::TILE);
    }
    for (int L = 0; L < levels; ++L) {
      const int shift = std::max(begin_bit, end_bit - 8 * (L + 1));
      const int nb = (end_bit - 8 * L) - shift;
      const uint32_t mask = (1u << nb) - 1u;
      Seg* cur = (L & 1) ? w.segs1 : w.segs0;
      Seg* nxt = (L & 1) ? w.segs0 : w.segs1;
      const int ib = in_buf(L), ob = out_buf(L);
      // B200_SEG_CONST (experimental): from level 1 on the histogram also keeps a per-segment OR / AND of the keys; level 0 is one
      // segment, covered by the key-range probe; the last level's buckets are final after its scatter anyway
      const bool segp = B200_SEG_CONST != 0 && L >= 1 && L + 1 < levels && w.seg_or != nullptr;

      { ProfScope prof("msb_sched", s); launch_k(level_prep_kernel, sms * 2, 512, 0, s, w.seg_hist, &ctr->num_segs[L], segp ? w.seg_or : nullptr, segp ? w.seg_and : nullptr); }
      TileHistArgs ha{};
      ha.keys = bufk[ib]; ha.descs = w.descs; ha.num_tiles_ptr = &ctr->num_tiles[L];
      ha.tile_off = w.tile_off; ha.group_tail = w.group_tail; ha.group_flag = w.group_flag; ha.seg_hist = w.seg_hist;
      ha.tile_cnt = w.tile_cnt;
      ha.shift = shift; ha.mask = mask; ha.tw_in = (L == 0) ? twid : 0; ha.tw = tw;
      ha.key_or = &ctr->key_or; ha.key_and = &ctr->key_and;
#if B200_HIST_TICKET
      ha.ticket = &ctr->part_ticket[L];
#endif
      const int hgrid = (int)std::min<uint32_t>(w.max_groups, (uint32_t)sms * 4);
      if (L == 0 && probe == 2) {            // exact per-key OR / AND (the first round found a single level-0 bucket)
        { ProfScope prof("tile_hist", s); launch_k(tile_hist_kernel<K, false, true>, hgrid, HIST_THREADS, 0, s, ha); }
        { ProfScope prof("msb_sched", s); launch_k(probe_eval_kernel<false>, 1, RADIX, 0, s, ctr, w.seg_hist, shift, begin_bit, end_bit); }
      } else if (segp) {
        ha.seg_or = w.seg_or; ha.seg_and = w.seg_and;
        { ProfScope prof("tile_hist", s); launch_k(tile_hist_kernel<K, false, false, B200_SEG_CONST != 0>, hgrid, HIST_THREADS, 0, s, ha); }
      } else {
        { ProfScope prof("tile_hist", s); launch_k(tile_hist_kernel<K, false, false>, hgrid, HIST_THREADS, 0, s, ha); }
        if (L == 0 && probe == 1) { ProfScope prof("msb_sched", s); launch_k(probe_eval_kernel<true>, 1, RADIX, 0, s, ctr, w.seg_hist, shift, begin_bit, end_bit); }
      }
      if (L == 0 && probe) {
        B200_CHECK(cudaMemcpyAsync(probe_buffer(), &ctr->key_or, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        B200_CHECK(cudaEventRecord(probe_event(), s));
      }
      { ProfScope prof("msb_sched", s); launch_k(group_carry_kernel, RADIX / 32, CARRY_WARPS * 32, 0, s, w.group_tail, w.group_flag, w.carry, &ctr->num_tiles[L]); }

      ClassifyArgs ca{};
      ca.segs = cur; ca.num_segs_ptr = &ctr->num_segs[L]; ca.seg_hist = w.seg_hist; ca.bins = w.bins;
      ca.next_segs = nxt; ca.num_next_ptr = &ctr->num_segs[L + 1]; ca.max_segs = w.max_segs;
      // buckets of this level: which on-chip algorithm.  More than 16 bits left -> one-shot counting sort; otherwise the unstable
      // keys-only engine still prefers it for its large unmerged buckets (one cell per key value: half the shared-memory traffic of
      // two LSD passes); merged runs, small buckets and the stable engine take the LSD kernels
      int list = (shift - begin_bit > 16) ? ALGO_COUNT : ALGO_LSD;
      const bool count_big = !ORDERED && VB == 0 && sizeof(K) == 4 && list == ALGO_LSD;
      const int big_list = count_big ? ALGO_COUNT : list;
      const bool big_special = list == ALGO_LSD && (use_rank || (count_big && use_bitmap));     // large unmerged buckets get their own list
      ca.locals = w.locals[big_list]; ca.num_locals_ptr = &ctr->num_locals[big_list]; ca.max_locals = w.max_locals;
      if (big_special) { ca.locals = w.locals[4]; ca.num_locals_ptr = &ctr->num_bitmap; }
      if (list == ALGO_LSD) { ca.locals_small = w.locals[3]; ca.num_small_ptr = &ctr->num_locals[2]; ca.small_cap = C::SMALL_CAP; }
      if (count_big || big_special) { ca.locals_merged = w.locals[ALGO_LSD]; ca.num_merged_ptr = &ctr->num_locals[ALGO_LSD]; }   // merged runs need LSD passes
      ca.error = &ctr->error; ca.shift = shift; ca.nb = nb; ca.last = (shift == begin_bit) ? 1 : 0;
      ca.local_cap = C::LOCAL_CAP; ca.merge_cap = (list == ALGO_LSD) ? (uint32_t)C::LOCAL_CAP : C::MERGE_CAP;
      // LSD levels: only small sub-buckets are merged (a merged run pays one more 8-bit pass over all of its keys); a bucket that
      // stands alone with at most 8 bits left is one counting pass of the LSD kernel, not a job for the one-shot kernels
      ca.merge_small = (list == ALGO_LSD) ? (uint32_t)C::LOCAL_CAP / 8u : ca.merge_cap;
      ca.dense_to_merged = (list == ALGO_LSD && shift - begin_bit <= 8) ? 1 : 0;
      // single buckets below the small configuration's capacity: with 9-16 bits left the rank kernel (one step) takes them from
      // B200_RANK_MIN keys on (measured); below that, and with at most 8 bits left, the small LSD configuration
      ca.small_max = C::SMALL_CAP;
      if (big_special && use_rank && !ca.dense_to_merged) {
        static const uint32_t rank_min = []() { const char* e = getenv("B200SORT_RANK_MIN"); return e ? (uint32_t)atoi(e) : (uint32_t)B200_RANK_MIN; }();
        ca.small_max = std::min<uint32_t>((uint32_t)C::SMALL_CAP, rank_min > 0 ? rank_min - 1 : 0u);
      }
      ca.out_buf = (uint32_t)ob;
      if (segp) { ca.seg_or = w.seg_or; ca.seg_and = w.seg_and; ca.locals_copy = w.locals[ALGO_LSD]; ca.num_copy_ptr = &ctr->num_locals[ALGO_LSD]; ca.begin_bit = begin_bit; }
      const int cgrid = (int)std::min<uint32_t>((w.max_segs + CLS_WARPS - 1) / CLS_WARPS, (uint32_t)sms * 4);
      { ProfScope prof("msb_sched", s); launch_k(classify_kernel, (L == 0 && segin == nullptr) ? 1 : cgrid, CLS_WARPS * 32, 0, s, ca); }

      ScatterArgs pa{};
      pa.keys_in = bufk[ib]; pa.keys_out = bufk[ob]; pa.vals_in = bufv[ib]; pa.vals_out = bufv[ob];
      pa.descs = w.descs; pa.num_tiles_ptr = &ctr->num_tiles[L];
      pa.bins = w.bins; pa.tile_off = w.tile_off; pa.carry = w.carry; pa.tile_cnt = w.tile_cnt;
      pa.shift = shift; pa.mask = mask; pa.tw_in = (L == 0) ? twid : 0; pa.tw_out = (shift == begin_bit) ? twid : 0; pa.tw = tw;
      B200_CHECK((launch_scatter<K, VB, MODE_SEG, ORDERED>(pa, w.max_tiles, s)));

      if (L + 1 < levels) {
        ProfScope prof("msb_sched", s);
        launch_k(scan_tiles_kernel, 1, SCAN_THREADS, 0, s, nxt, &ctr->num_segs[L + 1], w.tile_base, &ctr->num_tiles[L + 1], w.max_tiles, &ctr->error, C::TILE);
        launch_k(fill_descs_kernel, sms * 2, 256, 0, s, nxt, w.tile_base, &ctr->num_segs[L + 1], &ctr->num_tiles[L + 1], w.descs, C::TILE);
      }
    }
    if (fin_out) *fin_out = fin;
    la.keys_final = bufk[fin]; la.vals_final = bufv[fin];
    if (segin != nullptr) {      // segments that fit on chip untouched: still in caller form, all of [begin_bit, end_bit) to sort
      la.tw_in = twid; la.max_items = segin->num_segments;
      la.items = segin->direct; la.num_items_ptr = &ctr->num_direct[0];
      B200_CHECK((launch_local<K, VB, ALGO_LSD, ORDERED>(la, segin->num_segments, s)));
      la.items = segin->direct_small; la.num_items_ptr = &ctr->num_direct[1];
      B200_CHECK((launch_local<K, VB, ALGO_LSD, ORDERED, true>(la, segin->num_segments, s)));
    }
    la.tw_in = 0; la.max_items = w.max_locals;
    la.items = w.locals[0]; la.num_items_ptr = &ctr->num_locals[0];
    B200_CHECK((launch_local<K, VB, ALGO_LSD, ORDERED>(la, w.max_locals, s)));
    la.items = w.locals[3]; la.num_items_ptr = &ctr->num_locals[2];          // small buckets: the 256-thread configuration
    B200_CHECK((launch_local<K, VB, ALGO_LSD, ORDERED, true>(la, w.max_locals, s)));
    {
      la.items = w.locals[ALGO_COUNT]; la.num_items_ptr = &ctr->num_locals[ALGO_COUNT];      // more than 16 bits left (or the legacy exact-cell mode)
      if (end_bit - begin_bit > 24 || (!ORDERED && VB == 0 && sizeof(K) == 4 && !use_rank && !use_bitmap))
        B200_CHECK((launch_local<K, VB, ALGO_COUNT, ORDERED>(la, w.max_locals, s)));
      la.items = w.locals[4]; la.num_items_ptr = &ctr->num_bitmap;                            // large buckets with <= 16 bits left
      if constexpr (sizeof(K) == 4) {
        if (use_rank) {
          // buckets whose keys crowd their cells (heavy duplicates, or a key space as dense as the input) are passed on, unread, to the
          // 4-bit-counter variant of the same kernel
          la.dense = w.locals[5]; la.num_dense_ptr = &ctr->num_dense;
          B200_CHECK((launch_rank<K, VB, ORDERED>(la, w.max_locals, s)));
          la.dense = nullptr; la.num_dense_ptr = nullptr;
          la.items = w.locals[5]; la.num_items_ptr = &ctr->num_dense;
          B200_CHECK((launch_rank<K, VB, ORDERED, true>(la, w.max_locals, s)));
        }
      }
      if constexpr (!ORDERED && VB == 0 && sizeof(K) == 4) {
        if (use_bitmap) B200_CHECK((launch_bitmap<K, VB>(la, w.max_locals, s)));
      }
      la.items = w.locals[2]; la.num_items_ptr = &ctr->num_overflow;       // buckets those kernels handed back (heavy duplicates): LSD passes instead
      B200_CHECK((launch_local<K, VB, ALGO_LSD, ORDERED, false, true>(la, w.max_locals, s)));
      la.items = w.locals[6]; la.num_items_ptr = &ctr->num_overflow_small;      // ... the small ones in the 256-thread configuration
      B200_CHECK((launch_local<K, VB, ALGO_LSD, ORDERED, true, true>(la, w.max_locals, s)));
    }
    if (!probe) break;
    B200_CHECK(cudaEventSynchronize(probe_event()));
    const unsigned long long* hp = probe_buffer();
    if (probe == 1 && (hp[2] & 0xFFFFFFFFull) != 0) { probe = 2; continue; }      // one level-0 bucket: look at the lower bits too
    probe = 0;
    unsigned long long diff = hp[0] ^ hp[1];                                  // bits on which the keys differ
    diff &= (end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1ull)) & ~((1ull << begin_bit) - 1ull);
    int new_end = begin_bit;
    while (new_end < end_bit && (diff >> new_end) != 0) ++new_end;           // highest differing bit + 1
    if (new_end >= end_bit) break;                                            // the optimistic sort stands
    // the enqueued sort found its lists emptied (probe_eval_kernel): sort again on the narrower window
    end_bit = new_end;
    levels = (end_bit - begin_bit + 7) / 8;
    if (fin_in < 0) fin = levels & 1;
    if (levels == 0) {                // all keys equal on the sorted bits: the input order is the (stable) answer
      if (fin_out) *fin_out = fin;
      if (fin != 0) {
        B200_CHECK(cudaMemcpyAsync(bufk[fin], bufk[0], n * sizeof(K), cudaMemcpyDeviceToDevice, s));
        if (VB) B200_CHECK(cudaMemcpyAsync(bufv[fin], bufv[0], n * sizeof(V), cudaMemcpyDeviceToDevice, s));
      }
      return cudaSuccess;
    }
  }
  return cudaGetLastError();
}

// ===============================================================================================================
// Stable sort behind the cub::DeviceRadixSort call shape.
// k0/v0 = current (input), k1/v1 = alternate.  *selector (out) = which one holds the result.
// allow_overwrite = 0: the input buffers are left untouched and the result is delivered in the alternate buffers
// (CUB's pointer overloads, device_radix_sort.cuh:147-179; needs a third buffer inside the temporary storage,
// dispatch_radix_sort.cuh:1099-1104).
// Engine: the stable MSD hybrid above (fewer sweeps than one-pass-per-digit LSD whenever the leading digits split
// the keys: hist + 2 scatters + 1 on-chip sweep for 2^28 uniform keys of ANY width).  The onesweep LSD engine
// (one up-front histogram, one look-back scatter per digit) is kept: B200SORT_LSB_ENGINE=onesweep selects it, and it
// serves n >= 2^32.  Both are stable, so their results are identical bit for bit.
// ===============================================================================================================
inline bool use_onesweep_engine() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200SORT_LSB_ENGINE"); v = (e && e[0] == 'o') ? 1 : 0; }
  return v == 1;
}

template <typename K, int VB>
cudaError_t lsb_sort_impl(void* d_temp, size_t* temp_bytes, void* k0, void* k1, void* v0, void* v1, int* selector,
                          uint64_t n, const Twiddle& tw, int begin_bit, int end_bit, int allow_overwrite, cudaStream_t s) {
  using C = Cfg<K, VB>;
  using V = typename ValType<VB>::type;
  constexpr int KEY_BITS = sizeof(K) * 8;
  if (begin_bit < 0) begin_bit = 0;
  if (end_bit > KEY_BITS) end_bit = KEY_BITS;
  const int passes = end_bit > begin_bit ? (end_bit - begin_bit + 7) / 8 : 0;
  const bool onesweep = use_onesweep_engine() || n >= (1ull << 32);
  const uint64_t portion = (MAX_PORTION / C::TILE) * C::TILE;
  const uint64_t max_tiles = (std::min<uint64_t>(n, portion) + C::TILE - 1) / C::TILE;
  const bool need_third = !allow_overwrite && passes > 1 && n > (uint64_t)C::LOCAL_CAP;

  Carver cv(d_temp);
  MsdWorkspace w{};
  unsigned long long* hist = nullptr; uint64_t* pbins = nullptr; uint32_t* tick_status = nullptr;
  if (onesweep) {
    hist = cv.take<unsigned long long>((size_t)MAX_PASSES * RADIX);
    pbins = cv.take<uint64_t>(2 * RADIX);
    tick_status = cv.take<uint32_t>(64 + max_tiles * RADIX);     // [ticket | pad | status...], one memset clears both
    w.ctr = cv.take<MsbCounters>(1);
    w.locals[0] = cv.take<LocalItem>(1);
    w.max_locals = 1;                      // (inputs that fit one CTA take the single-item on-chip sort)
  } else {
    msd_carve<K, VB>(cv, n, w);
  }
  K* k2 = need_third ? cv.take<K>(n) : nullptr;
  V* v2 = (need_third && VB) ? cv.take<V>(n) : nullptr;
  if (d_temp == nullptr) { *temp_bytes = std::max<size_t>(cv.total(), 256); return cudaSuccess; }
  if (*temp_bytes < cv.total()) return cudaErrorInvalidValue;

  if (selector) *selector = 0;
  if (n == 0) return cudaSuccess;
  if (passes == 0) {
    if (!allow_overwrite) {
      B200_CHECK(cudaMemcpyAsync(k1, k0, n * sizeof(K), cudaMemcpyDeviceToDevice, s));
      if (VB) B200_CHECK(cudaMemcpyAsync(v1, v0, n * sizeof(V), cudaMemcpyDeviceToDevice, s));
      if (selector) *selector = 1;
    }
    return cudaSuccess;
  }
  if (!onesweep || n <= (uint64_t)C::LOCAL_CAP) {
    void* bufk[3] = {k0, k1, k2}; void* bufv[3] = {v0, v1, v2};
    // two buffers: the last possible level lands in buffer (passes & 1), so that is where everything is finished
    // (DoubleBuffer semantics: the selector says where); pointer overloads always deliver into the alternate buffers
    int fin = 1;
    // keys-only over the WHOLE key: equal keys are indistinguishable, so the cheaper unstable engine returns the identical result.
    // On a bit sub-range keys that tie on the window still differ elsewhere and must keep their input order (cub::DeviceRadixSort::SortKeys).
    const bool any_order = VB == 0 && begin_bit == 0 && end_bit == KEY_BITS;
    const cudaError_t e = any_order ? msd_sort_run<K, VB, false>(w, bufk, bufv, need_third ? 3 : 2, allow_overwrite ? -1 : 1, &fin, n, tw, begin_bit, end_bit, s)
                                    : msd_sort_run<K, VB, true>(w, bufk, bufv, need_third ? 3 : 2, allow_overwrite ? -1 : 1, &fin, n, tw, begin_bit, end_bit, s);
    if (selector) *selector = fin;
    return e;
  }

  // ---- onesweep LSD engine: all digit histograms in one read, then one look-back scatter launch per digit
  B200_CHECK(cudaMemsetAsync(hist, 0, (size_t)passes * RADIX * sizeof(unsigned long long), s));
  {
    HistAllArgs ha{};
    ha.keys = k0; ha.n = n; ha.num_passes = passes; ha.begin_bit = begin_bit; ha.end_bit = end_bit;
    ha.tw_in = 1; ha.tw = tw; ha.hist = hist;
    const int grid = (int)std::min<uint64_t>((uint64_t)num_sms() * 4, (n + 4095) / 4096);
    { ProfScope prof("hist_all", s); hist_all_kernel<K><<<grid, HIST_THREADS, 0, s>>>(ha); note_launch(); }
    { ProfScope prof("scan_bins", s); scan_bins_kernel<<<passes, RADIX, 0, s>>>(hist, 0); note_launch(); }
  }
  const void* src_k = k0; const void* src_v = v0;
  for (int p = 0; p < passes; ++p) {
    void* dst_k; void* dst_v;
    if (allow_overwrite) { dst_k = (p & 1) ? k0 : k1; dst_v = (p & 1) ? v0 : v1; }
    else { const bool to_out = ((passes - 1 - p) & 1) == 0; dst_k = to_out ? k1 : (void*)k2; dst_v = to_out ? v1 : (void*)v2; }
    const int shift = begin_bit + 8 * p;
    const int nb = end_bit - shift < 8 ? end_bit - shift : 8;
    int q = 0;
    for (uint64_t base = 0; base < n; base += portion, ++q) {
      const uint64_t pn = std::min<uint64_t>(portion, n - base);
      const uint32_t tiles = (uint32_t)((pn + C::TILE - 1) / C::TILE);
      B200_CHECK(cudaMemsetAsync(tick_status, 0, (64 + (size_t)tiles * RADIX) * sizeof(uint32_t), s));
      ScatterArgs pa{};
      pa.keys_in = src_k; pa.keys_out = dst_k; pa.vals_in = src_v; pa.vals_out = dst_v;
      pa.descs = nullptr; pa.num_tiles_ptr = nullptr;
      pa.num_tiles = tiles; pa.base = base; pa.n = pn;
      pa.bins = (q == 0) ? reinterpret_cast<uint64_t*>(hist + (size_t)p * RADIX) : pbins + ((q - 1) & 1) * RADIX;
      pa.bins_next = (base + pn < n) ? pbins + (q & 1) * RADIX : nullptr;
      pa.status = tick_status + 64; pa.ticket = tick_status;
      pa.shift = shift; pa.mask = (1u << nb) - 1u;
      pa.tw_in = (p == 0); pa.tw_out = (p == passes - 1); pa.tw = tw;
      B200_CHECK((launch_scatter<K, VB, MODE_LSB, true>(pa, tiles, s)));
    }
    src_k = dst_k; src_v = dst_v;
  }
  if (selector) *selector = allow_overwrite ? (passes & 1) : 1;
  return cudaSuccess;
}

// ===============================================================================================================
// Segmented stable sort behind the cub::DeviceSegmentedRadixSort call shape (lsb/cub/cub/device/device_segmented_radix_sort.cuh:
// 140-760; CUB runs one CTA-group per segment, dispatch_radix_sort.cuh:321-436).  Here the segments are simply the level-0
// buckets of the MSD engine: small segments go straight to the on-chip sorts (thousands per launch), large ones are split by
// the level loop -- no per-segment launch, no host read-back.  Buffer / selector / temporary-storage conventions as lsb_sort_impl.
// Elements that belong to no segment are unspecified in the output buffer (as in CUB).
// ===============================================================================================================
template <typename K, int VB>
cudaError_t segmented_sort_impl(void* d_temp, size_t* temp_bytes, void* k0, void* k1, void* v0, void* v1, int* selector,
                                uint64_t n, uint32_t num_segments, const void* d_begin, const void* d_end, int offset_bytes,
                                const Twiddle& tw, int begin_bit, int end_bit, int allow_overwrite, cudaStream_t s) {
  using C = Cfg<K, VB>;
  using V = typename ValType<VB>::type;
  constexpr int KEY_BITS = sizeof(K) * 8;
  if (begin_bit < 0) begin_bit = 0;
  if (end_bit > KEY_BITS) end_bit = KEY_BITS;
  if (n >= (1ull << 32) || (offset_bytes != 4 && offset_bytes != 8)) return cudaErrorInvalidValue;
  const int passes = end_bit > begin_bit ? (end_bit - begin_bit + 7) / 8 : 0;
  const bool overwrite = (allow_overwrite & 1) != 0;
  const bool need_third = !overwrite && passes > 1 && n > (uint64_t)C::LOCAL_CAP;

  Carver cv(d_temp);
  MsdWorkspace w{};
  msd_carve<K, VB>(cv, n, w);
  SegInput si{};
  si.begin = d_begin; si.end = d_end; si.num_segments = num_segments; si.offset_bytes = offset_bytes;
  si.direct = cv.take<LocalItem>((size_t)num_segments + 1);
  si.direct_small = cv.take<LocalItem>((size_t)num_segments + 1);
  K* k2 = need_third ? cv.take<K>(n) : nullptr;
  V* v2 = (need_third && VB) ? cv.take<V>(n) : nullptr;
  if (d_temp == nullptr) { *temp_bytes = std::max<size_t>(cv.total(), 256); return cudaSuccess; }
  if (*temp_bytes < cv.total()) return cudaErrorInvalidValue;

  if (selector) *selector = 0;
  if (n == 0 || num_segments == 0) return cudaSuccess;
  if (d_begin == nullptr || d_end == nullptr) return cudaErrorInvalidValue;
  if (passes == 0) {
    if (!overwrite) {
      B200_CHECK(cudaMemcpyAsync(k1, k0, n * sizeof(K), cudaMemcpyDeviceToDevice, s));
      if (VB) B200_CHECK(cudaMemcpyAsync(v1, v0, n * sizeof(V), cudaMemcpyDeviceToDevice, s));
      if (selector) *selector = 1;
    }
    return cudaSuccess;
  }
  void* bufk[3] = {k0, k1, k2}; void* bufv[3] = {v0, v1, v2};
  int fin = 1;
  // (see lsb_sort_impl; allow_overwrite bit 1: the caller vouches that keys which tie on the window are equal -- e.g. every segment
  // shares its bits above end_bit, like the buckets the multi-GPU exchange delivers)
  const bool any_order = VB == 0 && ((begin_bit == 0 && end_bit == KEY_BITS) || (allow_overwrite & 2));
  const cudaError_t e = any_order ? msd_sort_run<K, VB, false>(w, bufk, bufv, need_third ? 3 : 2, overwrite ? -1 : 1, &fin, n, tw, begin_bit, end_bit, s, &si)
                                  : msd_sort_run<K, VB, true>(w, bufk, bufv, need_third ? 3 : 2, overwrite ? -1 : 1, &fin, n, tw, begin_bit, end_bit, s, &si);
  if (selector) *selector = fin;
  return e;
}

// ===============================================================================================================
// Unstable MSB hybrid sort.  Both buffer pairs are clobbered; *out_keys / *out_vals = the buffers holding the result
// (the input buffers for 4- and 8-byte keys, like the reference: gpu_radix_sort.h:359-360).
// ===============================================================================================================
template <typename K, int VB>
cudaError_t msb_sort_impl(void* keys, void* vals, uint64_t n, void* keys_alt, void* vals_alt, const Twiddle& tw,
                          void* d_ws, size_t* ws_bytes, cudaStream_t s, void** out_keys, void** out_vals, int begin_bit, int end_bit) {
  constexpr int KEY_BITS = sizeof(K) * 8;
  if (begin_bit < 0) begin_bit = 0;
  if (end_bit > KEY_BITS) end_bit = KEY_BITS;
  if (end_bit < begin_bit) end_bit = begin_bit;
  const int LEVELS = (end_bit - begin_bit + 7) / 8;
  if (out_keys) *out_keys = keys;
  if (out_vals) *out_vals = vals;

  if (n >= (1ull << 32)) {     // beyond the 32-bit tile offsets: the stable LSD engine gives a valid result
    int sel = 0;
    cudaError_t e = lsb_sort_impl<K, VB>(d_ws, ws_bytes, keys, keys_alt, vals, vals_alt, &sel, n, tw, begin_bit, end_bit, 1, s);
    if (d_ws && e == cudaSuccess && sel) { if (out_keys) *out_keys = keys_alt; if (out_vals) *out_vals = vals_alt; }
    return e;
  }
  Carver cv(d_ws);
  MsdWorkspace w{};
  msd_carve<K, VB>(cv, n, w);
  if (d_ws == nullptr) { *ws_bytes = std::max<size_t>(cv.total(), 256); return cudaSuccess; }
  if (*ws_bytes < cv.total()) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  void* bufk[3] = {keys, keys_alt, nullptr}; void* bufv[3] = {vals, vals_alt, nullptr};
  if (LEVELS == 0) return cudaSuccess;
  // 4 / 8 levels: the last level lands in the input buffer, like the reference; fewer levels (bit sub-range, or leading bits
  // found constant by the key-range probe) may leave the result in the alternate buffers -- out_keys / out_vals say where
  int fin = 0;
  const cudaError_t e = msd_sort_run<K, VB, false>(w, bufk, bufv, 2, -1, &fin, n, tw, begin_bit, end_bit, s);
  if (out_keys) *out_keys = bufk[fin];
  if (out_vals) *out_vals = bufv[fin];
  return e;
}

// ===============================================================================================================
// Multi-GPU send partition: stable split of (keys, values) into `num_parts` contiguous parts by key range.
// The part of a key = number of splitters <= its top-`bits` bucket (transformed key).  d_local_counts is this
// rank's own top-`bits` histogram (b200_msd_histogram), from which the part sizes follow without another read.
// ===============================================================================================================
static __global__ void __launch_bounds__(256) part_offsets_kernel(const uint64_t* counts, int nbuckets, const uint32_t* splitters, int num_parts,
                                                                 uint64_t* part_offsets, uint64_t* bins, const uint64_t* dst_base) {
  __shared__ unsigned long long sums[MAX_PARTS];
  __shared__ uint32_t sp[MAX_PARTS];
  if (threadIdx.x < MAX_PARTS) { sums[threadIdx.x] = 0; sp[threadIdx.x] = (int)threadIdx.x < num_parts - 1 ? splitters[threadIdx.x] : 0xFFFFFFFFu; }
  __syncthreads();
  for (int b = threadIdx.x; b < nbuckets; b += blockDim.x) {
    const uint64_t c = counts[b];
    if (c) {
      int d = 0;
      for (int j = 0; j < num_parts - 1; ++j) d += ((uint32_t)b >= sp[j]) ? 1 : 0;
      atomicAdd(&sums[d], (unsigned long long)c);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t run = 0;
    for (int d = 0; d < RADIX; ++d) {
      bins[d] = (dst_base != nullptr) ? (d < num_parts ? dst_base[d] : 0ull) : run;       // own buffer per destination: caller's base
      if (d <= num_parts) part_offsets[d] = run;
      if (d < num_parts) run += sums[d];
    }
  }
}

template <typename K, int VB>
cudaError_t range_partition_impl(void* d_temp, size_t* temp_bytes, const void* kin, const void* vin, void* kout, void* vout,
                                 uint64_t n, const Twiddle& tw, int bits, const uint32_t* d_splitters, int num_parts,
                                 const uint64_t* d_local_counts, uint64_t* d_part_offsets, const uint64_t* d_dst_keys, const uint64_t* d_dst_vals,
                                 const uint64_t* d_dst_base, cudaStream_t s) {
  using C = Cfg<K, VB>;
  constexpr int KEY_BITS = sizeof(K) * 8;
  if (num_parts < 1 || num_parts > MAX_PARTS || bits < 1 || bits > 16 || n >= (1ull << 32)) return cudaErrorInvalidValue;
  const uint32_t max_tiles = (uint32_t)(n / C::TILE) + 2;
  const uint32_t max_groups = max_tiles / HIST_GROUP + 1;
  Carver cv(d_temp);
  uint64_t* bins = cv.take<uint64_t>(RADIX);
  MsbCounters* ctr = cv.take<MsbCounters>(1);
  Seg* seg = cv.take<Seg>(2);
  uint32_t* tile_base = cv.take<uint32_t>(4);
  TileDesc* descs = cv.take<TileDesc>(max_tiles);
  uint32_t* seg_hist = cv.take<uint32_t>(RADIX);
  uint32_t* tile_off = cv.take<uint32_t>((size_t)max_tiles * RADIX);
  uint32_t* group_tail = cv.take<uint32_t>((size_t)max_groups * RADIX);
  uint32_t* group_flag = cv.take<uint32_t>(max_groups);
  uint32_t* carry = cv.take<uint32_t>((size_t)max_groups * RADIX);
  if (d_temp == nullptr) { *temp_bytes = std::max<size_t>(cv.total(), 256); return cudaSuccess; }
  if (*temp_bytes < cv.total()) return cudaErrorInvalidValue;
  // part sizes follow from the caller's own top-bits histogram; bins[d] = start of part d (in the shared output, or 0-based
  // inside its own destination buffer + the caller's base when every part has its own buffer)
  part_offsets_kernel<<<1, 256, 0, s>>>(d_local_counts, 1 << bits, d_splitters, num_parts, d_part_offsets, bins, d_dst_keys ? d_dst_base : nullptr); note_launch();
  if (n == 0) return cudaGetLastError();
  const int sms = num_sms();
  B200_CHECK(cudaMemsetAsync(ctr, 0, sizeof(MsbCounters), s));
  B200_CHECK(cudaMemsetAsync(seg_hist, 0, RADIX * sizeof(uint32_t), s));
  msb_init_kernel<<<1, 32, 0, s>>>(seg, ctr, n); note_launch();
  scan_tiles_kernel<<<1, SCAN_THREADS, 0, s>>>(seg, &ctr->num_segs[0], tile_base, &ctr->num_tiles[0], max_tiles, &ctr->error, C::TILE); note_launch();
  fill_descs_kernel<<<sms * 2, 256, 0, s>>>(seg, tile_base, &ctr->num_segs[0], &ctr->num_tiles[0], descs, C::TILE); note_launch();
  TileHistArgs ha{};
  ha.keys = kin; ha.descs = descs; ha.num_tiles_ptr = &ctr->num_tiles[0];
  ha.tile_off = tile_off; ha.group_tail = group_tail; ha.group_flag = group_flag; ha.seg_hist = seg_hist;
  ha.shift = KEY_BITS - bits; ha.mask = 0xFFu; ha.tw_in = 1; ha.tw = tw; ha.splitters = d_splitters; ha.num_parts = num_parts;
  { ProfScope prof("tile_hist", s); tile_hist_kernel<K, true, false><<<(int)std::min<uint32_t>(max_groups, (uint32_t)sms * 4), HIST_THREADS, 0, s>>>(ha); note_launch(); }
  group_carry_kernel<<<RADIX / 32, CARRY_WARPS * 32, 0, s>>>(group_tail, group_flag, carry, &ctr->num_tiles[0]); note_launch();
  ScatterArgs pa{};
  pa.keys_in = kin; pa.keys_out = kout; pa.vals_in = vin; pa.vals_out = vout;
  pa.descs = descs; pa.num_tiles_ptr = &ctr->num_tiles[0];
  pa.bins = bins; pa.tile_off = tile_off; pa.carry = carry;
  pa.shift = KEY_BITS - bits; pa.mask = (uint32_t)(num_parts - 1);
  pa.tw_in = 1; pa.tw_out = 1; pa.tw = tw;
  pa.splitters = d_splitters; pa.num_parts = num_parts; pa.dst_keys = d_dst_keys; pa.dst_vals = d_dst_vals;
  B200_CHECK((launch_scatter<K, VB, MODE_RANGE, true>(pa, max_tiles, s)));
  return cudaGetLastError();
}

// ===============================================================================================================
// Multi-GPU exchange as level 0 of the sort (gpu_sort_b200/dist.py, ExchangeSorter).
// Every rank counts the leading 8-bit digit of its keys per tile (exchange_hist: ONE read, the same per-tile histogram pass as a
// sort level) and publishes its 256 digit totals; from the gathered G x 256 matrix every rank derives the same plan: digits are
// dealt to the G ranks in contiguous, balanced groups, and inside a destination's receive buffer the digits lie in ascending
// order, each digit's keys in source-rank order (stable).  exchange_scatter then runs the stable scatter of a sort level whose
// per-digit destinations are the PEERS' receive buffers (stores over NVLink from the scatter's coalesced write-out): after it,
// rank D holds level-0 buckets of the global sort, already split, and finishes them with a segmented sort on the remaining
// bits (b200_segmented_sort with the segment bounds the plan wrote) -- the exchange costs one sweep and replaces one.
// ===============================================================================================================
struct ExchangeWorkspace {
  MsbCounters* ctr; Seg* seg; uint32_t* tile_base; TileDesc* descs; uint32_t* seg_hist; uint32_t* tile_off; uint16_t* tile_cnt;
  uint32_t* group_tail; uint32_t* group_flag; uint32_t* carry; uint64_t* bins; uint8_t* digit_dest; uint32_t max_tiles, max_groups;
};
template <typename K, int VB>
inline void exchange_carve(Carver& cv, uint64_t n, ExchangeWorkspace& w) {
  using C = Cfg<K, VB>;
  w.max_tiles = (uint32_t)(n / C::TILE) + 2;
  w.max_groups = w.max_tiles / HIST_GROUP + 1;
  w.ctr = cv.take<MsbCounters>(1);
  w.seg = cv.take<Seg>(2);
  w.tile_base = cv.take<uint32_t>(4);
  w.descs = cv.take<TileDesc>(w.max_tiles);
  w.seg_hist = cv.take<uint32_t>(RADIX);
  w.tile_off = cv.take<uint32_t>((size_t)w.max_tiles * RADIX);
  w.tile_cnt = cv.take<uint16_t>((size_t)w.max_tiles * RADIX);
  w.group_tail = cv.take<uint32_t>((size_t)w.max_groups * RADIX);
  w.group_flag = cv.take<uint32_t>(w.max_groups);
  w.carry = cv.take<uint32_t>((size_t)w.max_groups * RADIX);
  w.bins = cv.take<uint64_t>(RADIX);
  w.digit_dest = cv.take<uint8_t>(RADIX);
}

// out[b] = keys of exchange bucket b = sum over the digits (b << xshift) .. ; entries past the last bucket are zero
static __global__ void __launch_bounds__(RADIX) widen_hist_kernel(const uint32_t* seg_hist, uint64_t* out, int xshift) {
  const unsigned b = threadIdx.x;
  unsigned long long sum = 0;
  if ((b << xshift) < (unsigned)RADIX)
    for (unsigned d = b << xshift; d < ((b + 1u) << xshift); ++d) sum += seg_hist[d];
  out[b] = sum;
}

// info: [0] keys this rank receives, [1] status (0 ok, 1 some rank would receive more than `cap`: nothing is exchanged),
//       [2] the largest receive count over all ranks, [3] first digit this rank owns, [4] one past the last digit it owns
static __global__ void __launch_bounds__(RADIX) exchange_plan_kernel(const uint64_t* matrix, int G, int rank, uint64_t cap, uint64_t* bins, uint8_t* digit_dest,
                                                                     uint64_t* seg_begin, uint64_t* seg_end, uint64_t* info, MsbCounters* ctr) {
  __shared__ unsigned long long cum[RADIX + 1];      // cum[d] = keys of all ranks with digit < d
  __shared__ uint32_t bound[MAX_PARTS + 1];          // rank j owns the digits [bound[j], bound[j + 1])
  const unsigned d = threadIdx.x;
  unsigned long long tot = 0, before = 0;
  for (int r = 0; r < G; ++r) { const unsigned long long c = matrix[(size_t)r * RADIX + d]; tot += c; if (r < rank) before += c; }
  cum[d + 1] = tot;
  if (d == 0) cum[0] = 0;
  __syncthreads();
  if (d == 0) for (int i = 1; i <= RADIX; ++i) cum[i] += cum[i - 1];
  __syncthreads();
  const unsigned long long n = cum[RADIX];
  if (d <= (unsigned)G) {
    // boundary j: the digit boundary whose cumulative count is closest to j * n / G (ties: the lower one); monotone by construction
    uint32_t b = d == 0 ? 0u : (d == (unsigned)G ? (uint32_t)RADIX : 0u);
    if (d > 0 && d < (unsigned)G) {
      const unsigned long long target = (unsigned long long)(((unsigned __int128)n * d) / (unsigned)G);
      uint32_t lo = 0, hi = RADIX;                   // first boundary with cum >= target
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (cum[mid] >= target) hi = mid; else lo = mid + 1; }
      b = lo;
      if (b > 0 && target - cum[b - 1] <= cum[b] - target) b -= 1;
    }
    bound[d] = b;
  }
  __syncthreads();
  if (d == 0) for (int j = 1; j <= G; ++j) if (bound[j] < bound[j - 1]) bound[j] = bound[j - 1];
  __syncthreads();
  uint32_t dest = 0;
  for (int j = 1; j < G; ++j) dest += bound[j] <= d ? 1u : 0u;
  const unsigned long long start = cum[d] - cum[bound[dest]];         // where digit d begins inside its destination's receive buffer
  bins[d] = start + before;                                           // ... and where this rank's share of it begins
  digit_dest[d] = (uint8_t)dest;
  const bool mine = dest == (uint32_t)rank;
  seg_begin[d] = mine ? start : 0ull;
  seg_end[d] = mine ? start + tot : 0ull;
  if (d == 0) {
    unsigned long long worst = 0;
    for (int j = 0; j < G; ++j) { const unsigned long long c = cum[bound[j + 1]] - cum[bound[j]]; worst = c > worst ? c : worst; }
    info[0] = cum[bound[rank + 1]] - cum[bound[rank]];
    info[1] = worst > cap ? 1ull : 0ull;
    info[2] = worst; info[3] = bound[rank]; info[4] = bound[rank + 1];
    if (worst > cap) ctr->num_tiles[0] = 0;                           // the scatter behind this kernel finds nothing to do
  }
}

template <typename K, int VB>
cudaError_t exchange_hist_impl(void* d_temp, size_t* temp_bytes, const void* kin, uint64_t n, const Twiddle& tw, int bucket_bits, uint64_t* d_hist, cudaStream_t s) {
  using C = Cfg<K, VB>;
  if (n >= (1ull << 32) || bucket_bits < 0 || bucket_bits > 8) return cudaErrorInvalidValue;
  const int digit_shift = (int)sizeof(K) * 8 - 8;        // ranking digit: the leading 8 bits; bucket = its leading bucket_bits bits
  Carver cv(d_temp);
  ExchangeWorkspace w{};
  exchange_carve<K, VB>(cv, n, w);
  if (d_temp == nullptr) { *temp_bytes = std::max<size_t>(cv.total(), 256); return cudaSuccess; }
  if (*temp_bytes < cv.total()) return cudaErrorInvalidValue;
  const int sms = num_sms();
  B200_CHECK(cudaMemsetAsync(w.ctr, 0, sizeof(MsbCounters), s));
  B200_CHECK(cudaMemsetAsync(w.seg_hist, 0, RADIX * sizeof(uint32_t), s));
  msb_init_kernel<<<1, 32, 0, s>>>(w.seg, w.ctr, n); note_launch();
  scan_tiles_kernel<<<1, SCAN_THREADS, 0, s>>>(w.seg, &w.ctr->num_segs[0], w.tile_base, &w.ctr->num_tiles[0], w.max_tiles, &w.ctr->error, C::TILE); note_launch();
  fill_descs_kernel<<<sms * 2, 256, 0, s>>>(w.seg, w.tile_base, &w.ctr->num_segs[0], &w.ctr->num_tiles[0], w.descs, C::TILE); note_launch();
  TileHistArgs ha{};
  ha.keys = kin; ha.descs = w.descs; ha.num_tiles_ptr = &w.ctr->num_tiles[0];
  ha.tile_off = w.tile_off; ha.group_tail = w.group_tail; ha.group_flag = w.group_flag; ha.seg_hist = w.seg_hist; ha.tile_cnt = w.tile_cnt;
  const int nb = std::min(8, (int)sizeof(K) * 8 - digit_shift);
  ha.shift = digit_shift; ha.mask = (1u << nb) - 1u; ha.tw_in = 1; ha.tw = tw;
  ha.ticket = &w.ctr->part_ticket[0];
  { ProfScope prof("tile_hist", s); tile_hist_kernel<K, false, false><<<(int)std::min<uint32_t>(w.max_groups, (uint32_t)sms * 4), HIST_THREADS, 0, s>>>(ha); note_launch(); }
  group_carry_kernel<<<RADIX / 32, CARRY_WARPS * 32, 0, s>>>(w.group_tail, w.group_flag, w.carry, &w.ctr->num_tiles[0]); note_launch();
  widen_hist_kernel<<<1, RADIX, 0, s>>>(w.seg_hist, d_hist, 8 - bucket_bits); note_launch();
  return cudaGetLastError();
}

template <typename K, int VB>
cudaError_t exchange_scatter_impl(void* d_temp, size_t* temp_bytes, const void* kin, const void* vin, uint64_t n, const Twiddle& tw, int bucket_bits,
                                  const uint64_t* d_matrix, int G, int rank, uint64_t cap, const uint64_t* d_dst_keys, const uint64_t* d_dst_vals,
                                  uint64_t* d_seg_begin, uint64_t* d_seg_end, uint64_t* d_info, cudaStream_t s) {
  using C = Cfg<K, VB>;
  if (n >= (1ull << 32) || G < 1 || G > MAX_PARTS || rank < 0 || rank >= G || cap >= (1ull << 32) || bucket_bits < 0 || bucket_bits > 8) return cudaErrorInvalidValue;
  const int digit_shift = (int)sizeof(K) * 8 - 8;
  Carver cv(d_temp);
  ExchangeWorkspace w{};
  exchange_carve<K, VB>(cv, n, w);
  if (d_temp == nullptr) { *temp_bytes = std::max<size_t>(cv.total(), 256); return cudaSuccess; }
  if (*temp_bytes < cv.total()) return cudaErrorInvalidValue;
  exchange_plan_kernel<<<1, RADIX, 0, s>>>(d_matrix, G, rank, cap, w.bins, w.digit_dest, d_seg_begin, d_seg_end, d_info, w.ctr); note_launch();
  if (n == 0) return cudaGetLastError();
  const int nb = std::min(8, (int)sizeof(K) * 8 - digit_shift);
  ScatterArgs pa{};
  pa.keys_in = kin; pa.vals_in = vin; pa.keys_out = nullptr; pa.vals_out = nullptr;
  pa.descs = w.descs; pa.num_tiles_ptr = &w.ctr->num_tiles[0];
  pa.bins = w.bins; pa.tile_off = w.tile_off; pa.carry = w.carry; pa.tile_cnt = w.tile_cnt;
  pa.shift = digit_shift; pa.mask = (1u << nb) - 1u; pa.tw_in = 1; pa.tw_out = 1; pa.tw = tw;
  pa.num_parts = G; pa.dst_keys = d_dst_keys; pa.dst_vals = d_dst_vals; pa.digit_dest = w.digit_dest; pa.xshift = 8 - bucket_bits;
  B200_CHECK((launch_scatter_exchange<K, VB>(pa, w.max_tiles, s)));
  return cudaGetLastError();
}

}  // namespace b200
