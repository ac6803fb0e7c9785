// include/shim/cub/device/device_radix_sort.cuh -- shadows <cub/device/device_radix_sort.cuh>.
//
// With `-I <repo>/include/shim` ahead of the CUB include path, the reference LSB driver (lsb/sort.cu:8:
// `#include <cub/device/device_radix_sort.cuh>`) compiles UNCHANGED against libb200sort.so: cub::DeviceRadixSort becomes
// the C-ABI-backed struct of include/b200sort_cub_shim.cuh; everything else (DoubleBuffer, CachingDeviceAllocator,
// CubDebugExit) still comes from the real CUB on the include path.
#pragma once
#include "../../../b200sort_cub_shim.cuh"
namespace cub { using DeviceRadixSort = B200DeviceRadixSort; }
