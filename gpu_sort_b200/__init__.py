"""gpu_sort_b200 -- host-side mirror of the reference's radix-sort entry points on top of the B200-native C ABI.

The product is ``libb200sort.so`` (hand-written sm_100a CUDA behind ``include/b200sort.h``).  This module is the thin
Python host layer used by the tests and the benchmark: it binds the C ABI with ctypes and keeps the reference's
names and argument meaning:

* ``DeviceRadixSort.SortPairs / SortKeys / SortPairsDescending / SortKeysDescending`` with a ``DoubleBuffer`` and the
  two-phase temporary-storage protocol   (reference: lsb/cub/cub/device/device_radix_sort.cuh:147-781, driver
  call shape lsb/sort.cu:25-76);
* ``rdxsrt_unstable_sort`` (device buffers, returns which buffers hold the result) and the host-pointer wrappers
  ``rdxsrt_unstable_sort_keys`` / ``rdxsrt_unstable_sort_pairs``  (reference: msb/src/sort/gpu_radix_sort.h:187-587).

torch is used for device memory and streams only.  There is NO CPU fallback: importing this module without the
built CUDA library raises.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SORT_LIB") or os.path.join(_HERE, "libb200sort.so")   # env override: kernel-variant A/B runs (tools/)

KEY_U32, KEY_U64, KEY_I32, KEY_I64, KEY_F32, KEY_F64 = range(6)
KEY_BYTES = {KEY_U32: 4, KEY_U64: 8, KEY_I32: 4, KEY_I64: 8, KEY_F32: 4, KEY_F64: 8}

_TORCH_KEY = {torch.int32: KEY_I32, torch.int64: KEY_I64, torch.float32: KEY_F32, torch.float64: KEY_F64}
for _name, _kt in (("uint32", KEY_U32), ("uint64", KEY_U64)):
    if hasattr(torch, _name):
        _TORCH_KEY[getattr(torch, _name)] = _kt
_NP_KEY = {np.dtype("uint32"): KEY_U32, np.dtype("uint64"): KEY_U64, np.dtype("int32"): KEY_I32,
           np.dtype("int64"): KEY_I64, np.dtype("float32"): KEY_F32, np.dtype("float64"): KEY_F64}


class B200SortError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C gpu_sort_b200/csrc`).  gpu_sort_b200 has no CPU or library fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, u64, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
    P = ctypes.POINTER
    lib.b200_version.restype = i32
    lib.b200_error_string.restype = ctypes.c_char_p
    lib.b200_error_string.argtypes = [i32]
    lib.b200_lsb_sort.restype = i32
    lib.b200_lsb_sort.argtypes = [vp, P(sz), vp, vp, vp, vp, P(i32), u64, i32, i32, i32, i32, i32, i32, vp]
    if hasattr(lib, "b200_segmented_sort") or not os.environ.get("B200SORT_LIB"):     # (older A/B variant builds lack it)
        lib.b200_segmented_sort.restype = i32
        lib.b200_segmented_sort.argtypes = [vp, P(sz), vp, vp, vp, vp, P(i32), u64, ctypes.c_uint32, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.b200_msb_sort.restype = i32
    lib.b200_msb_sort.argtypes = [vp, vp, u64, vp, vp, i32, i32, vp, P(sz), vp, P(vp), P(vp)]
    lib.b200_msb_sort_bits.restype = i32
    lib.b200_msb_sort_bits.argtypes = [vp, vp, u64, vp, vp, i32, i32, i32, i32, vp, P(sz), vp, P(vp), P(vp)]
    lib.b200_range_partition_to.restype = i32
    lib.b200_range_partition_to.argtypes = [vp, P(sz), vp, vp, u64, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp]
    if hasattr(lib, "b200_exchange_hist") or not os.environ.get("B200SORT_LIB"):
        lib.b200_exchange_hist.restype = i32
        lib.b200_exchange_hist.argtypes = [vp, P(sz), vp, u64, i32, i32, i32, vp, vp]
        lib.b200_exchange_scatter.restype = i32
        lib.b200_exchange_scatter.argtypes = [vp, P(sz), vp, vp, u64, i32, i32, i32, vp, i32, i32, u64, vp, vp, vp, vp, vp, vp]
    if hasattr(lib, "b200_sort_status"):
        lib.b200_sort_status.restype = i32
        lib.b200_sort_status.argtypes = [vp, vp, P(i32)]
        lib.b200_set_key_range_probe.restype = i32
        lib.b200_set_key_range_probe.argtypes = [i32]
        lib.b200_host_cache_release.restype = i32
    lib.b200_msb_sort_host.restype = i32
    lib.b200_msb_sort_host.argtypes = [vp, vp, u64, vp, vp, i32, i32]
    lib.b200_lsb_sort_host.restype = i32
    lib.b200_lsb_sort_host.argtypes = [vp, vp, u64, vp, vp, i32, i32, i32]
    lib.b200_msd_histogram.restype = i32
    lib.b200_msd_histogram.argtypes = [vp, u64, i32, i32, vp, vp]
    lib.b200_range_partition.restype = i32
    lib.b200_range_partition.argtypes = [vp, P(sz), vp, vp, vp, vp, u64, i32, i32, i32, vp, i32, vp, vp, vp]
    lib.b200_util_generate_keys.restype = i32
    lib.b200_util_generate_keys.argtypes = [vp, u64, u64, u64, i32, u64, i32, u64, vp]
    lib.b200_util_iota.restype = i32
    lib.b200_util_iota.argtypes = [vp, u64, u64, i32, vp]
    lib.b200_prof_enable.restype = i32
    lib.b200_prof_enable.argtypes = [i32]
    if hasattr(lib, "b200_prof_launches"):
        lib.b200_prof_launches.restype = ctypes.c_ulonglong
    lib.b200_prof_report.restype = i32
    lib.b200_prof_report.argtypes = [ctypes.c_char_p, sz]
    lib.b200_util_check.restype = i32
    lib.b200_util_check.argtypes = [vp, vp, u64, i32, i32, i32, vp, vp]
    return lib


lib = _load()


def _check(err: int, what: str):
    if err != 0:
        raise B200SortError(f"{what} failed: cudaError {err} ({lib.b200_error_string(err).decode()})")


def _ptr(t) -> ctypes.c_void_p:
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        return ctypes.c_void_p(t.ctypes.data)
    return ctypes.c_void_p(int(t))


def _stream(stream) -> ctypes.c_void_p:
    if stream is None:
        stream = torch.cuda.current_stream()
    return ctypes.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))


def key_type_of(t, key_type: Optional[int] = None) -> int:
    if key_type is not None:
        return key_type
    if isinstance(t, torch.Tensor):
        return _TORCH_KEY[t.dtype]
    return _NP_KEY[t.dtype]


def _value_bytes(v) -> int:
    if v is None:
        return 0
    b = v.element_size() if isinstance(v, torch.Tensor) else v.dtype.itemsize
    if b not in (4, 8):
        raise ValueError("values must be 4 or 8 bytes wide")
    return b


# ----------------------------------------------------------------------------------------------------------------
# LSB: cub::DeviceRadixSort call shape
# ----------------------------------------------------------------------------------------------------------------
class DoubleBuffer:
    """cub::DoubleBuffer (lsb/cub/cub/util_type.cuh:785): two device buffers and a selector."""

    def __init__(self, current: torch.Tensor, alternate: torch.Tensor):
        self.d_buffers = [current, alternate]
        self.selector = 0

    def Current(self) -> torch.Tensor:
        return self.d_buffers[self.selector]

    def Alternate(self) -> torch.Tensor:
        return self.d_buffers[self.selector ^ 1]


class DeviceRadixSort:
    """Mirror of cub::DeviceRadixSort (lsb/cub/cub/device/device_radix_sort.cuh).

    ``d_temp_storage=None`` returns the number of temporary bytes needed and does no work (two-phase protocol);
    otherwise the sort is enqueued on ``stream`` and the DoubleBuffer selectors are updated.
    """

    @staticmethod
    def _run(d_temp_storage, d_keys, d_values, num_items, begin_bit, end_bit, descending, stream, key_type,
             keys_out=None, values_out=None):
        overwrite = keys_out is None
        if overwrite:
            k_cur, k_alt = d_keys.Current(), d_keys.Alternate()
            v_cur, v_alt = (d_values.Current(), d_values.Alternate()) if d_values is not None else (None, None)
        else:
            k_cur, k_alt, v_cur, v_alt = d_keys, keys_out, d_values, values_out
        kt = key_type_of(k_cur, key_type)
        vb = _value_bytes(v_cur)
        if end_bit is None:
            end_bit = KEY_BYTES[kt] * 8
        nbytes = ctypes.c_size_t(0 if d_temp_storage is None else d_temp_storage.numel() * d_temp_storage.element_size())
        sel = ctypes.c_int(0)
        err = lib.b200_lsb_sort(_ptr(d_temp_storage), ctypes.byref(nbytes), _ptr(k_cur), _ptr(k_alt), _ptr(v_cur), _ptr(v_alt),
                                ctypes.byref(sel), num_items, kt, vb, begin_bit, end_bit, int(descending), int(overwrite),
                                _stream(stream))
        _check(err, "b200_lsb_sort")
        if d_temp_storage is None:
            return nbytes.value
        if overwrite:
            d_keys.selector ^= sel.value
            if d_values is not None:
                d_values.selector ^= sel.value
        return nbytes.value

    @staticmethod
    def SortPairs(d_temp_storage, d_keys, d_values, num_items, begin_bit=0, end_bit=None, stream=None, key_type=None,
                  d_keys_out=None, d_values_out=None):
        return DeviceRadixSort._run(d_temp_storage, d_keys, d_values, num_items, begin_bit, end_bit, False, stream, key_type,
                                    d_keys_out, d_values_out)

    @staticmethod
    def SortPairsDescending(d_temp_storage, d_keys, d_values, num_items, begin_bit=0, end_bit=None, stream=None,
                            key_type=None, d_keys_out=None, d_values_out=None):
        return DeviceRadixSort._run(d_temp_storage, d_keys, d_values, num_items, begin_bit, end_bit, True, stream, key_type,
                                    d_keys_out, d_values_out)

    @staticmethod
    def SortKeys(d_temp_storage, d_keys, num_items, begin_bit=0, end_bit=None, stream=None, key_type=None, d_keys_out=None):
        return DeviceRadixSort._run(d_temp_storage, d_keys, None, num_items, begin_bit, end_bit, False, stream, key_type,
                                    d_keys_out, None)

    @staticmethod
    def SortKeysDescending(d_temp_storage, d_keys, num_items, begin_bit=0, end_bit=None, stream=None, key_type=None,
                           d_keys_out=None):
        return DeviceRadixSort._run(d_temp_storage, d_keys, None, num_items, begin_bit, end_bit, True, stream, key_type,
                                    d_keys_out, None)


class DeviceSegmentedRadixSort:
    """Mirror of cub::DeviceSegmentedRadixSort (lsb/cub/cub/device/device_segmented_radix_sort.cuh:140-844): every segment
    ``[d_begin_offsets[i], d_end_offsets[i])`` is sorted on its own (stably), all segments in one call.  Offsets are int32 or
    int64 device tensors; the CSR form passes ``offsets[:-1]`` and ``offsets[1:]``.  Protocol as DeviceRadixSort."""

    @staticmethod
    def _run(d_temp_storage, d_keys, d_values, num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit, end_bit,
             descending, stream, key_type, keys_out=None, values_out=None, ties_are_equal=False):
        overwrite = keys_out is None
        if overwrite:
            k_cur, k_alt = d_keys.Current(), d_keys.Alternate()
            v_cur, v_alt = (d_values.Current(), d_values.Alternate()) if d_values is not None else (None, None)
        else:
            k_cur, k_alt, v_cur, v_alt = d_keys, keys_out, d_values, values_out
        kt = key_type_of(k_cur, key_type)
        vb = _value_bytes(v_cur)
        if end_bit is None:
            end_bit = KEY_BYTES[kt] * 8
        ob = d_begin_offsets.element_size() if d_begin_offsets is not None else 4
        if d_begin_offsets is not None and (d_end_offsets.element_size() != ob or ob not in (4, 8)):
            raise ValueError("segment offsets must both be int32 or both int64")
        nbytes = ctypes.c_size_t(0 if d_temp_storage is None else d_temp_storage.numel() * d_temp_storage.element_size())
        sel = ctypes.c_int(0)
        err = lib.b200_segmented_sort(_ptr(d_temp_storage), ctypes.byref(nbytes), _ptr(k_cur), _ptr(k_alt), _ptr(v_cur), _ptr(v_alt),
                                      ctypes.byref(sel), num_items, num_segments, _ptr(d_begin_offsets), _ptr(d_end_offsets), ob,
                                      kt, vb, begin_bit, end_bit, int(descending), int(overwrite) | (2 if ties_are_equal else 0), _stream(stream))
        _check(err, "b200_segmented_sort")
        if d_temp_storage is None:
            return nbytes.value
        if overwrite:
            d_keys.selector ^= sel.value
            if d_values is not None:
                d_values.selector ^= sel.value
        return nbytes.value

    @staticmethod
    def SortPairs(d_temp_storage, d_keys, d_values, num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit=0,
                  end_bit=None, stream=None, key_type=None, d_keys_out=None, d_values_out=None):
        return DeviceSegmentedRadixSort._run(d_temp_storage, d_keys, d_values, num_items, num_segments, d_begin_offsets, d_end_offsets,
                                             begin_bit, end_bit, False, stream, key_type, d_keys_out, d_values_out)

    @staticmethod
    def SortPairsDescending(d_temp_storage, d_keys, d_values, num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit=0,
                            end_bit=None, stream=None, key_type=None, d_keys_out=None, d_values_out=None):
        return DeviceSegmentedRadixSort._run(d_temp_storage, d_keys, d_values, num_items, num_segments, d_begin_offsets, d_end_offsets,
                                             begin_bit, end_bit, True, stream, key_type, d_keys_out, d_values_out)

    @staticmethod
    def SortKeys(d_temp_storage, d_keys, num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit=0, end_bit=None,
                 stream=None, key_type=None, d_keys_out=None):
        return DeviceSegmentedRadixSort._run(d_temp_storage, d_keys, None, num_items, num_segments, d_begin_offsets, d_end_offsets,
                                             begin_bit, end_bit, False, stream, key_type, d_keys_out, None)

    @staticmethod
    def SortKeysDescending(d_temp_storage, d_keys, num_items, num_segments, d_begin_offsets, d_end_offsets, begin_bit=0, end_bit=None,
                           stream=None, key_type=None, d_keys_out=None):
        return DeviceSegmentedRadixSort._run(d_temp_storage, d_keys, None, num_items, num_segments, d_begin_offsets, d_end_offsets,
                                             begin_bit, end_bit, True, stream, key_type, d_keys_out, None)


# ----------------------------------------------------------------------------------------------------------------
# MSB: rdxsrt_unstable_sort call shape
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class RDXSRT_SortedSequence:
    """msb/src/sort/gpu_radix_sort.h:169-184: the buffers that hold the sorted result."""
    sorted_keys: torch.Tensor
    sorted_values: Optional[torch.Tensor]


def rdxsrt_workspace_bytes(num_items: int, key_type: int, value_bytes: int) -> int:
    """Size of the optional pre-allocated workspace (the reference's pre_allocated_dm, gpu_radix_sort.h:89-141)."""
    nbytes = ctypes.c_size_t(0)
    _check(lib.b200_msb_sort(None, None, num_items, None, None, key_type, value_bytes, None, ctypes.byref(nbytes), None, None, None),
           "b200_msb_sort(size query)")
    return nbytes.value


def rdxsrt_unstable_sort(dev_keys: torch.Tensor, dev_values: Optional[torch.Tensor], key_count: int,
                         dev_sorted_keys_out: torch.Tensor, dev_sorted_values_out: Optional[torch.Tensor],
                         workspace: Optional[torch.Tensor] = None, stream=None, key_type: Optional[int] = None,
                         begin_bit: int = 0, end_bit: int = 64) -> RDXSRT_SortedSequence:
    """rdxsrt_unstable_sort<KeyT,ValueT,IndexT> (msb/src/sort/gpu_radix_sort.h:187-507).  Both buffer pairs are clobbered.
    begin_bit / end_bit (b200_msb_sort_bits) restrict the comparison to those bits of the transformed key."""
    kt = key_type_of(dev_keys, key_type)
    vb = _value_bytes(dev_values)
    ok, ov = ctypes.c_void_p(0), ctypes.c_void_p(0)
    if workspace is None:
        err = lib.b200_msb_sort_bits(_ptr(dev_keys), _ptr(dev_values), key_count, _ptr(dev_sorted_keys_out), _ptr(dev_sorted_values_out),
                                     kt, vb, begin_bit, end_bit, None, None, _stream(stream), ctypes.byref(ok), ctypes.byref(ov))
    else:
        nbytes = ctypes.c_size_t(workspace.numel() * workspace.element_size())
        err = lib.b200_msb_sort_bits(_ptr(dev_keys), _ptr(dev_values), key_count, _ptr(dev_sorted_keys_out), _ptr(dev_sorted_values_out),
                                     kt, vb, begin_bit, end_bit, _ptr(workspace), ctypes.byref(nbytes), _stream(stream), ctypes.byref(ok), ctypes.byref(ov))
    _check(err, "b200_msb_sort")
    keys = dev_keys if ok.value == dev_keys.data_ptr() or key_count == 0 else dev_sorted_keys_out
    vals = None
    if dev_values is not None:
        vals = dev_values if ov.value == dev_values.data_ptr() or key_count == 0 else dev_sorted_values_out
    return RDXSRT_SortedSequence(keys, vals)


def rdxsrt_unstable_sort_keys(keys: np.ndarray, sorted_keys_out: Optional[np.ndarray] = None, key_type: Optional[int] = None) -> np.ndarray:
    """Host-pointer wrapper, msb/src/sort/gpu_radix_sort.h:510-541 (H2D + sort + D2H inside the call)."""
    keys = np.ascontiguousarray(keys)
    out = np.empty_like(keys) if sorted_keys_out is None else sorted_keys_out
    _check(lib.b200_msb_sort_host(_ptr(keys), None, keys.size, _ptr(out), None, key_type_of(keys, key_type), 0), "b200_msb_sort_host")
    return out


def rdxsrt_unstable_sort_pairs(keys: np.ndarray, values: np.ndarray, sorted_keys_out=None, sorted_values_out=None,
                               key_type: Optional[int] = None):
    """Host-pointer wrapper, msb/src/sort/gpu_radix_sort.h:543-587."""
    keys = np.ascontiguousarray(keys); values = np.ascontiguousarray(values)
    ko = np.empty_like(keys) if sorted_keys_out is None else sorted_keys_out
    vo = np.empty_like(values) if sorted_values_out is None else sorted_values_out
    _check(lib.b200_msb_sort_host(_ptr(keys), _ptr(values), keys.size, _ptr(ko), _ptr(vo), key_type_of(keys, key_type), _value_bytes(values)),
           "b200_msb_sort_host")
    return ko, vo


def lsb_sort_host(keys: np.ndarray, values: Optional[np.ndarray] = None, descending: bool = False, key_type: Optional[int] = None):
    """Host-pointer stable LSB sort (same shape as the MSB wrappers)."""
    keys = np.ascontiguousarray(keys)
    ko = np.empty_like(keys)
    vo = None
    if values is not None:
        values = np.ascontiguousarray(values); vo = np.empty_like(values)
    _check(lib.b200_lsb_sort_host(_ptr(keys), _ptr(values), keys.size, _ptr(ko), _ptr(vo), key_type_of(keys, key_type), _value_bytes(values),
                                  int(descending)), "b200_lsb_sort_host")
    return ko, vo


# ----------------------------------------------------------------------------------------------------------------
# Device-side utilities (synthetic inputs, result checks)
# ----------------------------------------------------------------------------------------------------------------
DIST = {"uniform": 0, "entropy": 1, "zipf_rank": 2, "zipf_hash": 3, "sorted": 4, "reverse": 5, "constant": 6}


def generate_keys(out: torch.Tensor, seed: int = 0, dist="uniform", param: int = 0, start: int = 0, total: Optional[int] = None, stream=None):
    n = out.numel()
    _check(lib.b200_util_generate_keys(_ptr(out), n, start, n if total is None else total, out.element_size() * 8, seed,
                                       DIST[dist] if isinstance(dist, str) else dist, param, _stream(stream)), "b200_util_generate_keys")
    return out


def iota(out: torch.Tensor, start: int = 0, stream=None):
    _check(lib.b200_util_iota(_ptr(out), out.numel(), start, out.element_size(), _stream(stream)), "b200_util_iota")
    return out


def check(keys: torch.Tensor, values: Optional[torch.Tensor] = None, descending=False, key_type: Optional[int] = None, stream=None):
    """Returns (digest_sum, digest_xor, out_of_order_pairs, descending_value_pairs_inside_equal_keys)."""
    out = torch.zeros(4, dtype=torch.int64, device=keys.device)
    _check(lib.b200_util_check(_ptr(keys), _ptr(values), keys.numel(), key_type_of(keys, key_type), _value_bytes(values), int(descending),
                               _ptr(out), _stream(stream)), "b200_util_check")
    r = out.cpu().numpy().view(np.uint64)
    return int(r[0]), int(r[1]), int(r[2]), int(r[3])


def prof_enable(on: bool = True):
    """Bracket every kernel launch of the library with CUDA events (b200_prof_enable)."""
    _check(lib.b200_prof_enable(int(on)), "b200_prof_enable")


def sort_status(d_temp: torch.Tensor, stream=None) -> int:
    """Device-side status word of the last sort that used `d_temp` (b200_sort_status); 0 = ok.  Synchronises."""
    st = ctypes.c_int(0)
    _check(lib.b200_sort_status(_ptr(d_temp), _stream(stream), ctypes.byref(st)), "b200_sort_status")
    return st.value


def set_key_range_probe(enable: bool) -> bool:
    """b200_set_key_range_probe: False = sort calls never wait on the host.  Returns the previous setting."""
    return bool(lib.b200_set_key_range_probe(int(enable)))


def prof_launches() -> int:
    """Kernels launched by the library since prof_enable(True) (b200_prof_launches)."""
    return int(lib.b200_prof_launches())


def prof_report() -> dict:
    """{kernel family: (launches, total ms)} since prof_enable / the last report (b200_prof_report); synchronises."""
    buf = ctypes.create_string_buffer(8192)
    _check(lib.b200_prof_report(buf, len(buf)), "b200_prof_report")
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out
