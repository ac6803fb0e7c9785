// util.cu -- device-side utilities of the C ABI: synthetic inputs (same generator as
// oracle/radix_oracle.c so CPU and GPU see identical keys), size-independent result checks for the full-size
// parity tests, and the top-bits histogram of the multi-GPU path.
#include <cuda_runtime.h>
#include "../../include/b200sort.h"
#include "common.cuh"

using namespace b200;

namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t stream64(uint64_t seed, uint64_t i) { return mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ull); }

__device__ __forceinline__ uint64_t gen_key(uint64_t i, uint64_t n, int key_bits, uint64_t seed, int dist, uint64_t param) {
  const uint64_t all = key_bits == 32 ? 0xFFFFFFFFull : ~0ull;
  uint64_t k;
  switch (dist) {
    case 1: {
      if (param == 0) return 0;
      k = stream64(seed, i);
      for (uint64_t j = 1; j < param; ++j) k &= stream64(seed + 17 * j, i);
      break;
    }
    case 2: case 3: {
      const uint64_t r1 = stream64(seed, i), r2 = stream64(seed + 17, i);
      const unsigned j = (unsigned)((r1 >> 32) % 20u);
      const uint64_t rank = (1ull << j) + (r2 & ((1ull << j) - 1));
      k = (dist == 2) ? rank : mix64(rank);
      break;
    }
    case 4: case 5: {
      const uint64_t idx = (dist == 4) ? i : (n - 1 - i);
      const uint64_t step = all / (n ? n : 1);
      k = idx * step + (step > 1 ? stream64(seed, idx) % step : 0);
      break;
    }
    case 6: k = mix64(seed); break;
    default: k = stream64(seed, i); break;
  }
  return k & all;
}

__global__ void gen_kernel(void* out, uint64_t n, uint64_t start, uint64_t total, int key_bits, uint64_t seed, int dist, uint64_t param) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t k = gen_key(start + i, total, key_bits, seed, dist, param);
    if (key_bits == 32) reinterpret_cast<uint32_t*>(out)[i] = (uint32_t)k; else reinterpret_cast<uint64_t*>(out)[i] = k;
  }
}

__global__ void iota_kernel(void* out, uint64_t n, uint64_t start, int value_bytes) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (value_bytes == 4) reinterpret_cast<uint32_t*>(out)[i] = (uint32_t)(start + i); else reinterpret_cast<uint64_t*>(out)[i] = start + i;
  }
}

template <typename K>
__global__ void check_kernel(const K* keys, const void* vals, uint64_t n, Twiddle tw, int value_bytes, unsigned long long* out) {
  unsigned long long sum = 0, x = 0, bad = 0, vbad = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const K k = keys[i];
    uint64_t v = 0;
    if (value_bytes == 4) v = reinterpret_cast<const uint32_t*>(vals)[i];
    else if (value_bytes == 8) v = reinterpret_cast<const uint64_t*>(vals)[i];
    const uint64_t h = mix64((uint64_t)k * 0x9E3779B97F4A7C15ull ^ mix64(v + 1));
    sum += h; x ^= h;
    if (i > 0) {
      const K kp = keys[i - 1];
      const K a = twiddle_in<K>(kp, tw), b = twiddle_in<K>(k, tw);
      if (a > b) ++bad;
      if (a == b && value_bytes) {
        uint64_t vp = value_bytes == 4 ? (uint64_t)reinterpret_cast<const uint32_t*>(vals)[i - 1] : reinterpret_cast<const uint64_t*>(vals)[i - 1];
        if (vp > v) ++vbad;
      }
    }
  }
  // warp reduce, then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_down_sync(0xffffffffu, sum, o);
    x ^= __shfl_down_sync(0xffffffffu, x, o);
    bad += __shfl_down_sync(0xffffffffu, bad, o);
    vbad += __shfl_down_sync(0xffffffffu, vbad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], sum); atomicXor(&out[1], x);
    if (bad) atomicAdd(&out[2], bad);
    if (vbad) atomicAdd(&out[3], vbad);
  }
}

template <typename K>
__global__ void __launch_bounds__(512) msd_hist_kernel(const K* keys, uint64_t n, Twiddle tw, int bits, unsigned long long* counts) {
  extern __shared__ uint32_t sh[];
  const int nb = 1 << bits;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int shift = (int)sizeof(K) * 8 - bits;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto count = [&](K k) { atomicAdd(&sh[(uint32_t)(twiddle_in<K>(k, tw) >> shift)], 1u); };
  // 16-byte vector loads, four in flight per thread, when the pointer allows it; scalar tail
  constexpr uint64_t VEC = 16 / sizeof(K);
  const uint64_t nvec = (reinterpret_cast<uintptr_t>(keys) & 15u) == 0 ? n / VEC : 0;
  const uint4* pv = reinterpret_cast<const uint4*>(keys);
  for (uint64_t i0 = t0; i0 < nvec; i0 += 4 * stride) {
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = i0 + u * stride < nvec ? pv[i0 + u * stride] : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * stride < nvec) {
        if (sizeof(K) == 4) { count((K)q[u].x); count((K)q[u].y); count((K)q[u].z); count((K)q[u].w); }
        else { count((K)(((uint64_t)q[u].y << 32) | q[u].x)); count((K)(((uint64_t)q[u].w << 32) | q[u].z)); }
      }
  }
  for (uint64_t i = nvec * VEC + t0; i < n; i += stride) count(keys[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&counts[i], (unsigned long long)c);
  }
}

bool key_tw(int key_type, int descending, Twiddle* tw, int* kb) {
  Twiddle t{0, 0, 0};
  switch (key_type) {
    case B200_KEY_U32: *kb = 4; break;
    case B200_KEY_U64: *kb = 8; break;
    case B200_KEY_I32: *kb = 4; t.sign_mask = 0x80000000ull; break;
    case B200_KEY_I64: *kb = 8; t.sign_mask = 0x8000000000000000ull; break;
    case B200_KEY_F32: *kb = 4; t.sign_mask = 0x80000000ull; t.float_mask = 0xFFFFFFFFull; break;
    case B200_KEY_F64: *kb = 8; t.sign_mask = 0x8000000000000000ull; t.float_mask = ~0ull; break;
    default: return false;
  }
  if (descending) t.flip_mask = (*kb == 4) ? 0xFFFFFFFFull : ~0ull;
  *tw = t;
  return true;
}

int grid_for(uint64_t n, int threads) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t want = (n + threads - 1) / threads;
  const uint64_t cap = (uint64_t)sms * 8;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

// Store-bandwidth probe for peer (NVLink) or local memory: writes `bytes` to dst in one of three patterns.
//   mode 0: every warp instruction stores 128 contiguous bytes (4 bytes per lane)
//   mode 1: every warp instruction stores 512 contiguous bytes (16 bytes per lane)
//   mode 2: 4-byte lanes, but consecutive `chunk`-byte pieces go to pseudo-random `chunk`-aligned places (a scatter's runs)
//   mode 3: 16-byte lanes, pieces of `chunk` bytes at pseudo-random places
//   mode 4: mode 2 with every piece shifted by 52 bytes: a warp's 128 bytes straddle two 128-byte lines (a scatter run at arbitrary alignment)
__global__ void store_probe_kernel(uint32_t* dst, uint64_t bytes, int mode, uint32_t chunk) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
  if (mode == 0) {
    for (uint64_t i = t; i < bytes / 4; i += nt) dst[i] = (uint32_t)i;
  } else if (mode == 1) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (uint64_t i = t; i < bytes / 16; i += nt) d4[i] = make_uint4((uint32_t)i, 1, 2, 3);
  } else {
    const uint64_t nchunks = bytes / chunk;            // power of two expected
    const uint32_t per = chunk / (mode != 3 ? 4 : 16);   // lanes per chunk
    const uint64_t units = bytes / (mode != 3 ? 4 : 16);
    for (uint64_t i = t; i < units; i += nt) {
      const uint64_t c = i / per, o = i % per;
      const uint64_t pc = (c * 0x9E3779B97F4A7C15ull >> 20) & (nchunks - 1);     // scrambled chunk index (bijective enough for a bandwidth probe)
      if (mode == 2) dst[pc * (chunk / 4) + o] = (uint32_t)i;
      else if (mode == 4) { if (pc + 1 < nchunks) dst[pc * (chunk / 4) + o + 13] = (uint32_t)i; }
      else reinterpret_cast<uint4*>(dst)[pc * (chunk / 16) + o] = make_uint4((uint32_t)i, 1, 2, 3);
    }
  }
}

}  // namespace

extern "C" {

B200_API int b200_util_store_probe(void* dst, uint64_t bytes, int mode, uint32_t chunk, int grid, int block, b200_stream_t stream) {
  store_probe_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<uint32_t*>(dst), bytes, mode, chunk);
  return (int)cudaGetLastError();
}


int b200_util_generate_keys(void* d_keys, uint64_t num_items, uint64_t start_index, uint64_t total_items, int key_bits,
                            uint64_t seed, int dist, uint64_t param, b200_stream_t stream) {
  if (key_bits != 32 && key_bits != 64) return (int)cudaErrorInvalidValue;
  if (num_items == 0) return 0;
  gen_kernel<<<grid_for(num_items, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_keys, num_items, start_index, total_items, key_bits, seed, dist, param);
  return (int)cudaGetLastError();
}

int b200_util_iota(void* d_values, uint64_t num_items, uint64_t start, int value_bytes, b200_stream_t stream) {
  if (value_bytes != 4 && value_bytes != 8) return (int)cudaErrorInvalidValue;
  if (num_items == 0) return 0;
  iota_kernel<<<grid_for(num_items, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_values, num_items, start, value_bytes);
  return (int)cudaGetLastError();
}

int b200_util_check(const void* d_keys, const void* d_values, uint64_t num_items, int key_type, int value_bytes, int descending,
                    uint64_t* d_out, b200_stream_t stream) {
  Twiddle tw; int kb;
  if (!key_tw(key_type, descending, &tw, &kb)) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(d_out, 0, 4 * sizeof(uint64_t), s);
  if (e != cudaSuccess) return (int)e;
  if (num_items == 0) return 0;
  const int g = grid_for(num_items, 256);
  if (kb == 4) check_kernel<uint32_t><<<g, 256, 0, s>>>(reinterpret_cast<const uint32_t*>(d_keys), d_values, num_items, tw, value_bytes, reinterpret_cast<unsigned long long*>(d_out));
  else check_kernel<uint64_t><<<g, 256, 0, s>>>(reinterpret_cast<const uint64_t*>(d_keys), d_values, num_items, tw, value_bytes, reinterpret_cast<unsigned long long*>(d_out));
  return (int)cudaGetLastError();
}

int b200_msd_histogram(const void* d_keys, uint64_t num_items, int key_type, int bits, uint64_t* d_counts, b200_stream_t stream) {
  Twiddle tw; int kb;
  if (!key_tw(key_type, 0, &tw, &kb) || bits < 1 || bits > 14) return (int)cudaErrorInvalidValue;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(d_counts, 0, sizeof(uint64_t) << bits, s);
  if (e != cudaSuccess) return (int)e;
  if (num_items == 0) return 0;
  const size_t smem = sizeof(uint32_t) << bits;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int g = sms * 4;
  if (kb == 4) {
    if ((e = cudaFuncSetAttribute(msd_hist_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)) != cudaSuccess) return (int)e;
    msd_hist_kernel<uint32_t><<<g, 512, smem, s>>>(reinterpret_cast<const uint32_t*>(d_keys), num_items, tw, bits, reinterpret_cast<unsigned long long*>(d_counts));
  } else {
    if ((e = cudaFuncSetAttribute(msd_hist_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)) != cudaSuccess) return (int)e;
    msd_hist_kernel<uint64_t><<<g, 512, smem, s>>>(reinterpret_cast<const uint64_t*>(d_keys), num_items, tw, bits, reinterpret_cast<unsigned long long*>(d_counts));
  }
  return (int)cudaGetLastError();
}

}  // extern "C"
