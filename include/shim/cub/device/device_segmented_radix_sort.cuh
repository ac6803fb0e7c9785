// include/shim/cub/device/device_segmented_radix_sort.cuh -- shadows <cub/device/device_segmented_radix_sort.cuh>.
//
// With `-I <repo>/include/shim` ahead of the CUB include path, code that calls cub::DeviceSegmentedRadixSort
// (lsb/cub/cub/device/device_segmented_radix_sort.cuh:140-844; the reference tree exercises it from
// lsb/cub/test/test_device_radix_sort.cu:241-358) compiles unchanged against libb200sort.so.
#pragma once
#include "../../../b200sort_cub_shim.cuh"
namespace cub { using DeviceSegmentedRadixSort = B200DeviceSegmentedRadixSort; }
