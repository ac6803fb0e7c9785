// ubench_rank.cu -- micro-benchmark that decides the in-tile digit-ranking primitive on sm_100a.
// Measures, full chip, the throughput (G digit-ranks / s) of the candidate ways to compute "how many
// earlier keys of my warp / block have my 8-bit digit":
//   ballot : 8 x VOTE + logic, warp-private counters (stable)
//   match  : MATCH.ANY, warp-private counters (stable)
//   atomor : shared-memory atomicOr match masks (stable)
//   atomw  : shared-memory atomicAdd with return on warp-private counters (unstable)
//   atomb  : shared-memory atomicAdd with return on block-shared counters (unstable)
//   red    : shared-memory atomicAdd, result unused (histogram only)
//   none   : digit generation only (overhead floor)
// for uniform-random digits and for a constant digit.  Results go to profiles/ (see tools/README).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ubench_rank ubench_rank.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;
constexpr int IPT = 16;

enum Mode { BALLOT = 0, MATCH = 1, ATOMOR = 2, ATOMW = 3, ATOMB = 4, RED = 5, NONE = 6, NMODES = 7 };
static const char* mode_name[NMODES] = {"ballot", "match", "atomor", "atomw", "atomb", "red", "none"};

__device__ __forceinline__ unsigned match8_ballot(unsigned d) {
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool p = (d >> b) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, p);
    peers &= p ? m : ~m;
  }
  return peers;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS) rank_kernel(unsigned* out, int iters, unsigned digit_mask, unsigned seed) {
  __shared__ unsigned cnt[WARPS * 256];
  __shared__ unsigned masks[WARPS * 256];
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (int i = threadIdx.x; i < WARPS * 256; i += THREADS) { cnt[i] = 0; masks[i] = 0; }
  __syncthreads();
  unsigned* wc = cnt + w * 256;
  unsigned* wm = masks + w * 256;
  unsigned x = seed ^ (blockIdx.x * THREADS + threadIdx.x) * 2654435761u;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      x = x * 1664525u + 1013904223u;
      const unsigned d = (x >> 24) & digit_mask;
      if (MODE == BALLOT || MODE == MATCH) {
        const unsigned peers = (MODE == BALLOT) ? match8_ballot(d) : __match_any_sync(0xffffffffu, d);
        const unsigned base = wc[d];
        __syncwarp();
        const unsigned below = __popc(peers & lt);
        if (below == 0) wc[d] = base + __popc(peers);
        __syncwarp();
        acc += base + below;
      } else if (MODE == ATOMOR) {
        atomicOr(&wm[d], 1u << lane);
        __syncwarp();
        const unsigned peers = wm[d];
        const unsigned base = wc[d];
        __syncwarp();
        const unsigned below = __popc(peers & lt);
        if (below == 0) { wc[d] = base + __popc(peers); wm[d] = 0; }
        __syncwarp();
        acc += base + below;
      } else if (MODE == ATOMW) {
        acc += atomicAdd(&wc[d], 1u);
      } else if (MODE == ATOMB) {
        acc += atomicAdd(&cnt[d], 1u);
      } else if (MODE == RED) {
        atomicAdd(&wc[d], 1u);
      } else {
        acc += d;
      }
    }
  }
  __syncthreads();
  if (MODE == RED) acc += wc[lane];
  out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

template <int MODE>
static void run(unsigned* d_out, int blocks, int iters, unsigned digit_mask, const char* dist) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rank_kernel<MODE><<<blocks, THREADS>>>(d_out, 4, digit_mask, 1u);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    rank_kernel<MODE><<<blocks, THREADS>>>(d_out, iters, digit_mask, 7u + r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  const double ranks = (double)blocks * THREADS * IPT * iters;
  printf("{\"ubench\":\"rank\",\"mode\":\"%s\",\"dist\":\"%s\",\"blocks\":%d,\"ms\":%.4f,\"granks_per_s\":%.2f,\"err\":%d}\n",
         mode_name[MODE], dist, blocks, best, ranks / best * 1e-6, (int)err);
  fflush(stdout);
}

int main(int argc, char** argv) {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int sm_clock_khz = 0; cudaDeviceGetAttribute(&sm_clock_khz, cudaDevAttrClockRate, dev);
  printf("{\"device\":\"%s\",\"sms\":%d,\"sm_clock_khz\":%d}\n", p.name, p.multiProcessorCount, sm_clock_khz);
  const int iters = argc > 1 ? atoi(argv[1]) : 400;
  unsigned* d_out; cudaMalloc(&d_out, (size_t)148 * 8 * THREADS * 4);
  for (int occ = 1; occ <= 4; occ *= 2) {
    const int blocks = p.multiProcessorCount * occ;
    run<BALLOT>(d_out, blocks, iters, 0xffu, "uniform");
    run<MATCH>(d_out, blocks, iters, 0xffu, "uniform");
    run<ATOMOR>(d_out, blocks, iters, 0xffu, "uniform");
    run<ATOMW>(d_out, blocks, iters, 0xffu, "uniform");
    run<ATOMB>(d_out, blocks, iters, 0xffu, "uniform");
    run<RED>(d_out, blocks, iters, 0xffu, "uniform");
    run<NONE>(d_out, blocks, iters, 0xffu, "uniform");
  }
  const int blocks = p.multiProcessorCount * 2;
  run<BALLOT>(d_out, blocks, iters, 0x0u, "constant");
  run<MATCH>(d_out, blocks, iters, 0x0u, "constant");
  run<ATOMOR>(d_out, blocks, iters, 0x0u, "constant");
  run<ATOMW>(d_out, blocks, iters, 0x0u, "constant");
  run<ATOMB>(d_out, blocks, iters, 0x0u, "constant");
  run<RED>(d_out, blocks, iters, 0x0u, "constant");
  run<BALLOT>(d_out, blocks, iters, 0x3u, "4values");
  run<MATCH>(d_out, blocks, iters, 0x3u, "4values");
  run<ATOMW>(d_out, blocks, iters, 0x3u, "4values");
  run<ATOMB>(d_out, blocks, iters, 0x3u, "4values");
  cudaFree(d_out);
  return 0;
}
