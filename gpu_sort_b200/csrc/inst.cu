// inst.cu -- one explicit instantiation unit per (key width, value width); compiled six times by the Makefile
// with -DB200_K=<uint32_t|uint64_t> -DB200_VB=<0|4|8> so the heavy kernels build in parallel.
#include "sort_impl.cuh"
#include "sort_api.h"

namespace b200 {
template cudaError_t lsb_sort_impl<B200_K, B200_VB>(void*, size_t*, void*, void*, void*, void*, int*, uint64_t, const Twiddle&, int, int, int, cudaStream_t);
template cudaError_t segmented_sort_impl<B200_K, B200_VB>(void*, size_t*, void*, void*, void*, void*, int*, uint64_t, uint32_t, const void*, const void*, int,
                                                          const Twiddle&, int, int, int, cudaStream_t);
template cudaError_t msb_sort_impl<B200_K, B200_VB>(void*, void*, uint64_t, void*, void*, const Twiddle&, void*, size_t*, cudaStream_t, void**, void**, int, int);
template cudaError_t range_partition_impl<B200_K, B200_VB>(void*, size_t*, const void*, const void*, void*, void*, uint64_t, const Twiddle&, int,
                                                           const uint32_t*, int, const uint64_t*, uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*, cudaStream_t);
template cudaError_t exchange_hist_impl<B200_K, B200_VB>(void*, size_t*, const void*, uint64_t, const Twiddle&, int, uint64_t*, cudaStream_t);
template cudaError_t exchange_scatter_impl<B200_K, B200_VB>(void*, size_t*, const void*, const void*, uint64_t, const Twiddle&, int, const uint64_t*, int, int, uint64_t,
                                                            const uint64_t*, const uint64_t*, uint64_t*, uint64_t*, uint64_t*, cudaStream_t);
}  // namespace b200
