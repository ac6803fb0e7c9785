"""ctypes binding of oracle/liboracle.so (CPU restatement of the reference's sorts).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by the product."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "liboracle.so")

KT = {"u32": 0, "u64": 1, "i32": 2, "i64": 3, "f32": 4, "f64": 5}
NP_OF = {"u32": np.uint32, "u64": np.uint64, "i32": np.int32, "i64": np.int64, "f32": np.float32, "f64": np.float64}
DIST = {"uniform": 0, "entropy": 1, "zipf_rank": 2, "zipf_hash": 3, "sorted": 4, "reverse": 5, "constant": 6}


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
        lib.oracle_lsb_sort.restype = i32
        lib.oracle_lsb_sort.argtypes = [vp, vp, u64, i32, i32, i32, i32, i32, vp, vp, i32]
        lib.oracle_msb_sort.restype = i32
        lib.oracle_msb_sort.argtypes = [vp, vp, u64, i32, i32, vp, vp, i32, i32]
        lib.oracle_segmented_sort.restype = i32
        lib.oracle_segmented_sort.argtypes = [vp, vp, u64, i32, i32, u64, vp, vp, i32, i32, i32, vp, vp]
        lib.oracle_gen_keys.restype = None
        lib.oracle_gen_keys.argtypes = [vp, u64, u64, u64, i32, u64, i32, u64]
        lib.oracle_digest.restype = None
        lib.oracle_digest.argtypes = [vp, vp, u64, i32, i32, ctypes.POINTER(u64), ctypes.POINTER(u64)]
        lib.oracle_count_unsorted.restype = u64
        lib.oracle_count_unsorted.argtypes = [vp, u64, i32, i32]
        lib.oracle_twiddle_in.restype = u64
        lib.oracle_twiddle_in.argtypes = [u64, i32]
        lib.oracle_twiddle_out.restype = u64
        lib.oracle_twiddle_out.argtypes = [u64, i32]

    @staticmethod
    def _p(a):
        return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)

    def gen_keys(self, n, key_bits=32, seed=0, dist="uniform", param=0, start=0, total=None):
        out = np.empty(n, dtype=np.uint32 if key_bits == 32 else np.uint64)
        self.lib.oracle_gen_keys(self._p(out), n, start, n if total is None else total, key_bits, seed, DIST[dist], param)
        return out

    def lsb_sort(self, keys, vals=None, key_type="u32", begin_bit=0, end_bit=None, descending=False, threads=0):
        keys = np.ascontiguousarray(keys)
        kb = keys.dtype.itemsize * 8
        ko = np.empty_like(keys)
        vo = np.empty_like(vals) if vals is not None else None
        rc = self.lib.oracle_lsb_sort(self._p(keys), self._p(vals), keys.size, KT[key_type], 0 if vals is None else vals.dtype.itemsize,
                                      begin_bit, kb if end_bit is None else end_bit, int(descending), self._p(ko), self._p(vo), threads)
        assert rc == 0
        return ko, vo

    def segmented_sort(self, keys, vals, begin, end, key_type="u32", begin_bit=0, end_bit=None, descending=False):
        """Every [begin[i], end[i]) sorted stably on its own; elements outside all segments copied through."""
        keys = np.ascontiguousarray(keys)
        begin = np.ascontiguousarray(begin, dtype=np.int64); end = np.ascontiguousarray(end, dtype=np.int64)
        kb = keys.dtype.itemsize * 8
        ko = np.empty_like(keys)
        vo = np.empty_like(vals) if vals is not None else None
        rc = self.lib.oracle_segmented_sort(self._p(keys), self._p(vals), keys.size, KT[key_type], 0 if vals is None else vals.dtype.itemsize,
                                            begin.size, self._p(begin), self._p(end), begin_bit, kb if end_bit is None else end_bit,
                                            int(descending), self._p(ko), self._p(vo))
        assert rc == 0
        return ko, vo

    def msb_sort(self, keys, vals=None, key_type="u32", local_cap=0, merge_thresh=-1):
        keys = np.ascontiguousarray(keys)
        ko = np.empty_like(keys)
        vo = np.empty_like(vals) if vals is not None else None
        rc = self.lib.oracle_msb_sort(self._p(keys), self._p(vals), keys.size, KT[key_type], 0 if vals is None else vals.dtype.itemsize,
                                      self._p(ko), self._p(vo), local_cap, merge_thresh)
        assert rc == 0
        return ko, vo

    def digest(self, keys, vals=None):
        s, x = ctypes.c_uint64(0), ctypes.c_uint64(0)
        self.lib.oracle_digest(self._p(keys), self._p(vals), keys.size, keys.dtype.itemsize * 8, 0 if vals is None else vals.dtype.itemsize,
                               ctypes.byref(s), ctypes.byref(x))
        return s.value, x.value

    def count_unsorted(self, keys, key_type="u32", descending=False):
        return self.lib.oracle_count_unsorted(self._p(keys), keys.size, KT[key_type], int(descending))


def load():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "radix_oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return Oracle(ctypes.CDLL(SO))
