// compat.h -- force-included glue (nvcc -include) so that the UNMODIFIED reference MSB sources under
// /root/reference/msb compile against the CUDA 12.9 toolkit's CUB 2.8.2 for sm_100a.
// Test infrastructure only (builds oracle/_ref/); never part of the product library.
// The reference uses cub::If / cub::Equals (removed from CUB 2.x; cuda_radix_sort.h:438,1362-1366,1504)
// and the pre-Volta __shfl_up (cuda_radix_sort_common.h:280,284).
#pragma once
#include <cub/cub.cuh>
#include <type_traits>
namespace cub {
template <bool C, typename A, typename B> struct If { typedef typename std::conditional<C, A, B>::type Type; };
template <typename A, typename B> struct Equals { enum { VALUE = std::is_same<A, B>::value ? 1 : 0, NEGATE = VALUE ? 0 : 1 }; };
}
#define __shfl_up(v, d) __shfl_up_sync(0xffffffffu, (v), (d))
