// async.cuh -- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier completion, the sm_90+/sm_100a way to stage
// a tile of keys into shared memory without tying up registers or issue slots: one thread arms the barrier with
// the byte count and launches the copy, everybody waits on the barrier's phase parity.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (follow with __syncthreads)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy; src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// orders earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory accesses through explicit 32-bit shared-window addresses.  The hot per-key paths keep their table bases in
// registers and add the element offset themselves: left to the compiler, the base of the dynamic shared-memory window was
// re-derived (S2R SR_CgaCtaId + LEA) for every key under register pressure (profiles/r02_rank_v2.txt).
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(uint32_t addr, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_or(uint32_t addr, uint32_t v) { uint32_t o; asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory"); return o; }
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) { uint32_t o; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory"); return o; }
__device__ __forceinline__ uint64_t lds_u64(uint32_t addr) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory"); return v; }
template <typename T> __device__ __forceinline__ T lds_t(uint32_t addr) { if (sizeof(T) == 4) return (T)lds_u32(addr); else return (T)lds_u64(addr); }
template <typename T> __device__ __forceinline__ void sts_t(uint32_t addr, T v) { if (sizeof(T) == 4) sts_u32(addr, (uint32_t)v); else sts_u64(addr, (uint64_t)v); }

// A [off, off+cnt) element window of a global array as a 16-byte aligned byte range for bulk_g2s:
// the copy starts `skew` elements before the window and may cover up to 15 bytes after it (same 16-byte granule as
// the last element, hence inside the same allocation granule as valid data).
template <typename T>
struct BulkWindow {
  const char* src; uint32_t bytes; uint32_t skew;
  __device__ __forceinline__ BulkWindow(const T* base, uint64_t off, uint32_t cnt) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(base + off);
    const uintptr_t a0 = a & ~(uintptr_t)15;
    skew = (uint32_t)((a - a0) / sizeof(T));
    bytes = (uint32_t)(((a - a0) + (uintptr_t)cnt * sizeof(T) + 15) & ~(uintptr_t)15);
    src = reinterpret_cast<const char*>(a0);
  }
};

}  // namespace b200
