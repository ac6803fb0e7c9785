// cpu_sort.cpp -- host-sort baselines named by BASELINE.json's north_star ("std::sort/parallel sort timed on the host's
// cores with the core count stated") and used by the reference's slow value check (msb/tests/test_sort_pairs.cu:80-109,
// the only place std::sort appears in the reference).  TEST / BENCH INFRASTRUCTURE ONLY: loaded by bench.py's
// cpu_baseline leg and tests/; never by the product.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <parallel/algorithm>
#include <omp.h>

template <typename T>
static double timed_sort(T* keys, uint64_t n, int threads) {
  auto t0 = std::chrono::steady_clock::now();
  if (threads == 1) std::sort(keys, keys + n);
  else {
    if (threads > 0) omp_set_num_threads(threads);
    __gnu_parallel::sort(keys, keys + n);
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

extern "C" {
// Sorts in place; returns seconds.  threads == 1 -> std::sort, otherwise __gnu_parallel::sort on `threads` cores (<=0: all).
double cpu_sort_u32(uint32_t* keys, uint64_t n, int threads) { return timed_sort(keys, n, threads); }
double cpu_sort_u64(uint64_t* keys, uint64_t n, int threads) { return timed_sort(keys, n, threads); }
int cpu_sort_max_threads(void) { return omp_get_max_threads(); }
}
