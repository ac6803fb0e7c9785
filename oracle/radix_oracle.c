/*
 * radix_oracle.c -- CPU restatement of the reference's two radix sorts.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (gpu_sort_b200/libb200sort.so) never links, loads or calls it.
 *
 * Parity pin: the reference stores no golden vectors -- its own tests compare against
 * an independent sorter (msb/tests/test_sort_keys.cpp:47-80 -> cub::DeviceRadixSort + memcmp;
 * lsb/cub/test/test_device_radix_sort.cu:554-611,634-696 -> std::stable_sort).  This oracle is therefore pinned
 * (a) against numpy's sort / stable argsort on the reference's own test families (tests/test_oracle.py), and
 * (b) on the GPU box against the UNMODIFIED reference compiled from /root/reference (oracle/_ref, built by
 *     oracle/Makefile), whose outputs on fixed seeds are also committed as digests under tests/golden/.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { KT_U32 = 0, KT_U64 = 1, KT_I32 = 2, KT_I64 = 3, KT_F32 = 4, KT_F64 = 5 };

static int key_bits_of(int key_type) { return (key_type == KT_U32 || key_type == KT_I32 || key_type == KT_F32) ? 32 : 64; }

/* ---- order-preserving bit transforms: lsb/cub/cub/util_type.cuh:966-974 (unsigned: identity),
 *      :1009-1017 (signed: flip the sign bit), :1079-1089 (floating point: negative -> ~bits, else flip sign). ---- */
uint64_t oracle_twiddle_in(uint64_t k, int key_type) {
  switch (key_type) {
    case KT_I32: return (k ^ 0x80000000ull) & 0xFFFFFFFFull;
    case KT_I64: return k ^ 0x8000000000000000ull;
    case KT_F32: return ((k & 0x80000000ull) ? ~k : (k ^ 0x80000000ull)) & 0xFFFFFFFFull;
    case KT_F64: return (k & 0x8000000000000000ull) ? ~k : (k ^ 0x8000000000000000ull);
    default: return k;
  }
}
uint64_t oracle_twiddle_out(uint64_t k, int key_type) {
  switch (key_type) {
    case KT_I32: return (k ^ 0x80000000ull) & 0xFFFFFFFFull;
    case KT_I64: return k ^ 0x8000000000000000ull;
    case KT_F32: return ((k & 0x80000000ull) ? (k ^ 0x80000000ull) : ~k) & 0xFFFFFFFFull;
    case KT_F64: return (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
    default: return k;
  }
}

static inline uint64_t load_key(const void* p, uint64_t i, int kb) {
  return kb == 32 ? (uint64_t)((const uint32_t*)p)[i] : ((const uint64_t*)p)[i];
}
static inline void store_key(void* p, uint64_t i, int kb, uint64_t v) {
  if (kb == 32) ((uint32_t*)p)[i] = (uint32_t)v; else ((uint64_t*)p)[i] = v;
}
static inline uint64_t load_val(const void* p, uint64_t i, int vb) {
  return vb == 4 ? (uint64_t)((const uint32_t*)p)[i] : ((const uint64_t*)p)[i];
}
static inline void store_val(void* p, uint64_t i, int vb, uint64_t v) {
  if (vb == 4) ((uint32_t*)p)[i] = (uint32_t)v; else ((uint64_t*)p)[i] = v;
}

/* ------------------------------------------------------------------------------------------------------------
 * Stable LSD radix sort = cub::DeviceRadixSort::{SortKeys,SortPairs}[Descending] over bits [begin_bit,end_bit)
 * (lsb/cub/cub/device/device_radix_sort.cuh:147-179,248-273,328-360,424-449).  Pass structure follows
 * DispatchRadixSort::InvokePasses / InvokePass (lsb/cub/cub/device/dispatch/dispatch_radix_sort.cuh:899-976,
 * 1050-1159): per pass   upsweep  = per-"block" digit histogram      (agent_radix_sort_upsweep.cuh:392-447),
 *                        scan     = exclusive scan of the digit-major spine (dispatch_radix_sort.cuh:120-146),
 *                        downsweep= stable rank + scatter             (agent_radix_sort_downsweep.cuh:484-580).
 * Here a "block" is one OpenMP thread owning an even share of the input (grid_even_share.cuh).  The digit width
 * does not change the result of a stable LSD sort, so 8-bit digits are used throughout (the reference plans
 * 6/7-bit passes on sm_5x, dispatch_radix_sort.cuh:1113-1118).  Descending order = sort ascending on the
 * complemented twiddled key (agent_radix_sort_downsweep.cuh uses the inverted digit for IS_DESCENDING), which
 * keeps equal keys in input order exactly like the reference.
 * keys_out/vals_out receive the result; inputs are not modified.  threads<=0 -> all cores.
 * ------------------------------------------------------------------------------------------------------------ */
int oracle_lsb_sort(const void* keys_in, const void* vals_in, uint64_t n, int key_type, int value_bytes,
                    int begin_bit, int end_bit, int descending, void* keys_out, void* vals_out, int threads) {
  const int kb = key_bits_of(key_type);
  if (end_bit > kb) end_bit = kb;
  if (begin_bit < 0) begin_bit = 0;
  if (value_bytes != 0 && value_bytes != 4 && value_bytes != 8) return -1;
  const uint64_t all = (kb == 32) ? 0xFFFFFFFFull : ~0ull;
  const uint64_t desc = descending ? all : 0;
  uint64_t* ka = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  uint64_t* kbuf = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  uint64_t *va = NULL, *vbuf = NULL;
  if (value_bytes) { va = (uint64_t*)malloc((n ? n : 1) * 8); vbuf = (uint64_t*)malloc((n ? n : 1) * 8); }
  if (!ka || !kbuf || (value_bytes && (!va || !vbuf))) return -2;
  int T = threads;
#ifdef _OPENMP
  if (T <= 0) T = omp_get_max_threads();
#else
  T = 1;
#endif
  if ((uint64_t)T > n / 4096 + 1) T = (int)(n / 4096 + 1);
#pragma omp parallel for num_threads(T) schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    ka[i] = (oracle_twiddle_in(load_key(keys_in, i, kb), key_type) ^ desc) & all;   /* TwiddleIn fused into the first load */
    if (value_bytes) va[i] = load_val(vals_in, i, value_bytes);
  }
  uint64_t* hist = (uint64_t*)malloc((size_t)T * 256 * sizeof(uint64_t));
  for (int bit = begin_bit; bit < end_bit; bit += 8) {
    const int nb = (end_bit - bit < 8) ? (end_bit - bit) : 8;
    const uint64_t mask = (1ull << nb) - 1;
    memset(hist, 0, (size_t)T * 256 * sizeof(uint64_t));
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
      const int t = omp_get_thread_num();
#else
      const int t = 0;
#endif
      const uint64_t lo = n * (uint64_t)t / T, hi = n * (uint64_t)(t + 1) / T;
      uint64_t* h = hist + (size_t)t * 256;
      for (uint64_t i = lo; i < hi; ++i) h[(ka[i] >> bit) & mask]++;            /* upsweep */
#pragma omp barrier
#pragma omp single
      {                                                                          /* spine scan, digit-major */
        uint64_t run = 0;
        for (int d = 0; d < 256; ++d)
          for (int tt = 0; tt < T; ++tt) { uint64_t c = hist[(size_t)tt * 256 + d]; hist[(size_t)tt * 256 + d] = run; run += c; }
      }
      for (uint64_t i = lo; i < hi; ++i) {                                       /* downsweep: stable scatter */
        const uint64_t k = ka[i];
        const uint64_t dst = h[(k >> bit) & mask]++;
        kbuf[dst] = k;
        if (value_bytes) vbuf[dst] = va[i];
      }
    }
    { uint64_t* tmp = ka; ka = kbuf; kbuf = tmp; tmp = va; va = vbuf; vbuf = tmp; }
  }
#pragma omp parallel for num_threads(T) schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    store_key(keys_out, i, kb, oracle_twiddle_out((ka[i] ^ desc) & all, key_type));  /* TwiddleOut fused into the last store */
    if (value_bytes) store_val(vals_out, i, value_bytes, va[i]);
  }
  free(hist); free(ka); free(kbuf); free(va); free(vbuf);
  return 0;
}

/*
 * Segmented stable sort: every segment [begin[i], end[i]) sorted on its own, the rest of the array copied through.
 * Follows the reference's own CPU solution for cub::DeviceSegmentedRadixSort, lsb/cub/test/test_device_radix_sort.cu:669-676
 * (per segment: std::stable_sort, descending = reverse + stable_sort + reverse, which keeps equal keys in input order), i.e.
 * oracle_lsb_sort applied to each slice.  offsets are int64 here.
 */
int oracle_segmented_sort(const void* keys_in, const void* vals_in, uint64_t n, int key_type, int value_bytes,
                          uint64_t num_segments, const int64_t* begin, const int64_t* end,
                          int begin_bit, int end_bit, int descending, void* keys_out, void* vals_out) {
  const int kbytes = key_bits_of(key_type) / 8;
  memcpy(keys_out, keys_in, n * (uint64_t)kbytes);
  if (value_bytes) memcpy(vals_out, vals_in, n * (uint64_t)value_bytes);
  for (uint64_t i = 0; i < num_segments; ++i) {
    if (end[i] <= begin[i]) continue;                       /* empty segment (device_segmented_radix_sort.cuh:150) */
    if (begin[i] < 0 || (uint64_t)end[i] > n) return -3;
    const uint64_t b = (uint64_t)begin[i], c = (uint64_t)(end[i] - begin[i]);
    const int rc = oracle_lsb_sort((const char*)keys_in + b * kbytes, value_bytes ? (const char*)vals_in + b * value_bytes : NULL, c, key_type,
                                   value_bytes, begin_bit, end_bit, descending, (char*)keys_out + b * kbytes,
                                   value_bytes ? (char*)vals_out + b * value_bytes : NULL, 1);
    if (rc) return rc;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Unstable MSB "hybrid radix sort" = rdxsrt_unstable_sort (msb/src/sort/gpu_radix_sort.h:187-507).
 *   pass loop, 8-bit digits from the most significant byte down      gpu_radix_sort.h:205,279-351,366-486
 *   per bucket: 256-bin histogram of digit (key >> (8k-8(pass+1)))&255  cuda_radix_sort.h:657-740 (:701)
 *   exclusive prefix sum -> sub-bucket offsets                        cuda_radix_sort.h:1014-1036
 *   classify sub-buckets: empty / tiny (merge neighbours while the running sum stays < merge threshold 3000,
 *     cuda_radix_sort.h:1079-1131,1185-1232; cuda_radix_sort_config.h:4) / local (<= largest local-sort
 *     KPB, gpu_sort_config.h:43-141) / non-local (gets another counting pass, cuda_radix_sort.h:1258-1269)
 *   counting-sort scatter of non-local buckets                        cuda_radix_sort.h:363-479
 *   local sort of small buckets on the remaining low bits (plus the current digit when merged),
 *     counting sort on the low 8 bits then a stable block radix sort on the rest    cuda_radix_sort.h:1332-1620
 * The GPU reference is unstable (chunk reservation by atomicAdd, cuda_radix_sort.h:408-417); this restatement is
 * one deterministic member of its output set: the key sequence is THE sorted sequence, the (key,value) multiset
 * is preserved; the order of values inside a run of equal keys is unspecified by the reference and here happens
 * to be input order.  local_cap<=0 selects the reference's largest local-sort KPB for the type pair.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct { uint64_t off, cnt; int pass; } seg_t;

static void local_sort(uint64_t* k, uint64_t* v, uint64_t* tk, uint64_t* tv, uint64_t cnt, int top_bit, int has_v) {
  /* stable LSD on bits [0, top_bit) */
  uint64_t h[257];
  for (int bit = 0; bit < top_bit; bit += 8) {
    const int nb = (top_bit - bit < 8) ? (top_bit - bit) : 8;
    const uint64_t mask = (1ull << nb) - 1;
    memset(h, 0, sizeof(h));
    for (uint64_t i = 0; i < cnt; ++i) h[((k[i] >> bit) & mask) + 1]++;
    for (int d = 0; d < 256; ++d) h[d + 1] += h[d];
    for (uint64_t i = 0; i < cnt; ++i) { uint64_t dst = h[(k[i] >> bit) & mask]++; tk[dst] = k[i]; if (has_v) tv[dst] = v[i]; }
    memcpy(k, tk, cnt * 8); if (has_v) memcpy(v, tv, cnt * 8);
  }
}

int oracle_msb_sort(const void* keys_in, const void* vals_in, uint64_t n, int key_type, int value_bytes,
                    void* keys_out, void* vals_out, int local_cap, int merge_thresh) {
  const int kb = key_bits_of(key_type);
  const int kbytes = kb / 8, has_v = value_bytes != 0;
  if (value_bytes != 0 && value_bytes != 4 && value_bytes != 8) return -1;
  if (local_cap <= 0) {                       /* gpu_sort_config.h:43-141: max KPT*TPB of the default config set */
    if (!has_v) local_cap = (kbytes == 4) ? 18 * 512 : 11 * 384;
    else if (kbytes == 4 && value_bytes == 4) local_cap = 15 * 384;
    else local_cap = 15 * 256;
  }
  if (merge_thresh < 0) merge_thresh = 3000;  /* cuda_radix_sort_config.h:4 RDXSRT_CFG_MERGE_LOCREC_THRESH */
  const int num_passes = kbytes;              /* gpu_radix_sort.h:205 */
  uint64_t* a = (uint64_t*)malloc((n ? n : 1) * 8);
  uint64_t* b = (uint64_t*)malloc((n ? n : 1) * 8);
  uint64_t *av = NULL, *bv = NULL;
  if (has_v) { av = (uint64_t*)malloc((n ? n : 1) * 8); bv = (uint64_t*)malloc((n ? n : 1) * 8); }
  for (uint64_t i = 0; i < n; ++i) {
    a[i] = oracle_twiddle_in(load_key(keys_in, i, kb), key_type);           /* TwiddleIn on pass 0, cuda_radix_sort.h:699-701 */
    if (has_v) av[i] = load_val(vals_in, i, value_bytes);
  }
  /* work list of non-local buckets for the current pass; data of pass p lives in `cur`, is scattered to `alt` */
  size_t cap_segs = 1024, nseg = 0, nnext = 0;
  seg_t* segs = (seg_t*)malloc(cap_segs * sizeof(seg_t));
  seg_t* next = (seg_t*)malloc(cap_segs * sizeof(seg_t));
  uint64_t* fin = (uint64_t*)malloc((n ? n : 1) * 8);      /* the "final" buffer local sorts write to, gpu_radix_sort.h:359-360,401 */
  uint64_t* finv = has_v ? (uint64_t*)malloc((n ? n : 1) * 8) : NULL;
  uint64_t* tk = (uint64_t*)malloc((size_t)(local_cap > 0 ? local_cap : 1) * 8 + 8 * (size_t)merge_thresh + 64);
  uint64_t* tv = (uint64_t*)malloc((size_t)(local_cap > 0 ? local_cap : 1) * 8 + 8 * (size_t)merge_thresh + 64);
  uint64_t *cur = a, *alt = b, *curv = av, *altv = bv;
  if (n > 0) { segs[0].off = 0; segs[0].cnt = n; segs[0].pass = 0; nseg = 1; }
  for (int pass = 0; pass < num_passes && nseg > 0; ++pass) {
    const int shift = kb - 8 * (pass + 1);
    nnext = 0;
    for (size_t s = 0; s < nseg; ++s) {
      const uint64_t off = segs[s].off, cnt = segs[s].cnt;
      uint64_t h[256], o[257];
      memset(h, 0, sizeof(h));
      for (uint64_t i = 0; i < cnt; ++i) h[(cur[off + i] >> shift) & 255]++;         /* histogram */
      o[0] = 0; for (int d = 0; d < 256; ++d) o[d + 1] = o[d] + h[d];                /* prefix sum */
      { uint64_t w[256]; memcpy(w, o, sizeof(w));
        for (uint64_t i = 0; i < cnt; ++i) {                                        /* partition */
          uint64_t dst = off + w[(cur[off + i] >> shift) & 255]++;
          alt[dst] = cur[off + i]; if (has_v) altv[dst] = curv[off + i];
        } }
      /* classify sub-buckets (cuda_radix_sort.h:1079-1131,1185-1269) */
      int d = 0;
      while (d < 256) {
        if (h[d] == 0) { ++d; continue; }
        if (shift == 0) {                                   /* last digit consumed: bucket is final, in `alt` */
          memcpy(fin + off + o[d], alt + off + o[d], h[d] * 8);
          if (has_v) memcpy(finv + off + o[d], altv + off + o[d], h[d] * 8);
          ++d; continue;
        }
        if (h[d] < (uint64_t)merge_thresh) {                /* tiny: merge following neighbours while sum < threshold */
          uint64_t sum = h[d]; int e = d + 1;
          while (e < 256 && sum + h[e] < (uint64_t)merge_thresh) { sum += h[e]; ++e; }
          const int merged = (e - d) > 1;
          uint64_t* pk = alt + off + o[d]; uint64_t* pv = has_v ? altv + off + o[d] : NULL;
          local_sort(pk, pv, tk, tv, sum, merged ? shift + 8 : shift, has_v);   /* merged buckets also sort the current digit */
          memcpy(fin + off + o[d], pk, sum * 8); if (has_v) memcpy(finv + off + o[d], pv, sum * 8);
          d = e; continue;
        }
        if (h[d] <= (uint64_t)local_cap) {                  /* local */
          uint64_t* pk = alt + off + o[d]; uint64_t* pv = has_v ? altv + off + o[d] : NULL;
          local_sort(pk, pv, tk, tv, h[d], shift, has_v);
          memcpy(fin + off + o[d], pk, h[d] * 8); if (has_v) memcpy(finv + off + o[d], pv, h[d] * 8);
          ++d; continue;
        }
        if (nnext == cap_segs) {                            /* non-local: another counting pass */
          cap_segs *= 2; segs = (seg_t*)realloc(segs, cap_segs * sizeof(seg_t)); next = (seg_t*)realloc(next, cap_segs * sizeof(seg_t));
        }
        next[nnext].off = off + o[d]; next[nnext].cnt = h[d]; next[nnext].pass = pass + 1; ++nnext; ++d;
      }
    }
    { seg_t* t = segs; segs = next; next = t; nseg = nnext; }
    { uint64_t* t = cur; cur = alt; alt = t; t = curv; curv = altv; altv = t; }
  }
  for (uint64_t i = 0; i < n; ++i) {
    store_key(keys_out, i, kb, oracle_twiddle_out(fin[i], key_type));          /* TwiddleOut on the last digit / local sort, :445,1606-1610 */
    if (has_v) store_val(vals_out, i, value_bytes, finv[i]);
  }
  free(a); free(b); free(av); free(bv); free(segs); free(next); free(fin); free(finv); free(tk); free(tv);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Synthetic inputs: portable counter-based generator, identical on CPU (here, numpy in
 * gpu_sort_b200/gen.py) and on the device (gpu_sort_b200/csrc/util.cu).
 *   stream(seed, i) = mix64(seed + (i+1) * 0x9E3779B97F4A7C15)      (splitmix64 finaliser)
 * dist 0 uniform | 1 AND of `param` streams seeded seed+17*j (the reference's entropy levels,
 * msb/tests/data_gen.h:44-76; param 0 -> all-zero keys) | 2 zipf-like, key = rank | 3 zipf-like, key = mix64(rank)
 * | 4 presorted ascending | 5 presorted descending | 6 constant (= mix64(seed)).
 * "zipf-like" = s~1 over 2^20 ranks, integer-only: octave j uniform in [0,20), rank = 2^j + (r & (2^j - 1)),
 * i.e. every octave of ranks carries equal mass as under a 1/x law.
 * ------------------------------------------------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t stream(uint64_t seed, uint64_t i) { return mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ull); }

uint64_t oracle_gen_key(uint64_t i, uint64_t n, int key_bits, uint64_t seed, int dist, uint64_t param) {
  const uint64_t all = key_bits == 32 ? 0xFFFFFFFFull : ~0ull;
  uint64_t k;
  switch (dist) {
    case 1: { if (param == 0) return 0; k = stream(seed, i); for (uint64_t j = 1; j < param; ++j) k &= stream(seed + 17 * j, i); break; }
    case 2: case 3: {
      const uint64_t r1 = stream(seed, i), r2 = stream(seed + 17, i);
      const unsigned j = (unsigned)((r1 >> 32) % 20u);
      const uint64_t rank = (1ull << j) + (r2 & ((1ull << j) - 1));
      k = (dist == 2) ? rank : mix64(rank);
      break; }
    case 4: case 5: {
      const uint64_t idx = (dist == 4) ? i : (n - 1 - i);
      const uint64_t step = (all / (n ? n : 1));           /* >= 1 while n <= 2^key_bits - 1 */
      k = idx * step + (step > 1 ? stream(seed, idx) % step : 0);
      break; }
    case 6: k = mix64(seed); break;
    default: k = stream(seed, i); break;
  }
  return k & all;
}
void oracle_gen_keys(void* out, uint64_t n, uint64_t start, uint64_t total_n, int key_bits, uint64_t seed, int dist, uint64_t param) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    uint64_t k = oracle_gen_key(start + (uint64_t)i, total_n, key_bits, seed, dist, param);
    if (key_bits == 32) ((uint32_t*)out)[i] = (uint32_t)k; else ((uint64_t*)out)[i] = k;
  }
}

/* Order-independent multiset digest of (key,value) pairs: sum and xor of mix64(key*GOLD ^ mix64(value+1)).
 * Used for the full-size property tests (same multiset before/after the sort). */
void oracle_digest(const void* keys, const void* vals, uint64_t n, int key_bits, int value_bytes, uint64_t* out_sum, uint64_t* out_xor) {
  uint64_t s = 0, x = 0;
#pragma omp parallel for reduction(+ : s) reduction(^ : x) schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    uint64_t k = load_key(keys, i, key_bits);
    uint64_t v = value_bytes ? load_val(vals, i, value_bytes) : 0;
    uint64_t h = mix64(k * 0x9E3779B97F4A7C15ull ^ mix64(v + 1));
    s += h; x ^= h;
  }
  *out_sum = s; *out_xor = x;
}

/* number of adjacent inversions under the twiddled (total) order; 0 <=> sorted */
uint64_t oracle_count_unsorted(const void* keys, uint64_t n, int key_type, int descending) {
  const int kb = key_bits_of(key_type);
  uint64_t bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (int64_t i = 1; i < (int64_t)n; ++i) {
    uint64_t a = oracle_twiddle_in(load_key(keys, i - 1, kb), key_type), b = oracle_twiddle_in(load_key(keys, i, kb), key_type);
    bad += descending ? (a < b) : (a > b);
  }
  return bad;
}
