"""CPU tests of the drop-in boundary: the C-ABI library loads and exports exactly what include/b200sort.h declares,
the product never touches the oracle, and argument errors come back as error codes (no compute is launched here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200sort.h")
LIB = os.path.join(ROOT, "gpu_sort_b200", "libb200sort.so")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("b200_lsb_sort", "b200_segmented_sort", "b200_msb_sort", "b200_msb_sort_host", "b200_lsb_sort_host", "b200_msd_histogram",
              "b200_range_partition", "b200_util_generate_keys", "b200_util_iota", "b200_util_check", "b200_version"):
        assert s in syms


@pytest.mark.skipif(not os.path.exists(LIB), reason="libb200sort.so not built (python -c 'import __graft_entry__ as g; g.build()')")
def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/b200sort.h but not exported"
    lib.b200_version.restype = ctypes.c_int
    assert lib.b200_version() >= 100


@pytest.mark.skipif(not os.path.exists(LIB), reason="libb200sort.so not built")
def test_library_exports_nothing_but_the_abi():
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    extra = {s for s in exported if not s.startswith("b200_") and not s.startswith("_")}
    assert not extra, f"unexpected exported symbols: {sorted(extra)[:5]}"
    assert set(declared_symbols()) <= exported


@pytest.mark.skipif(not os.path.exists(LIB), reason="libb200sort.so not built")
def test_size_queries_and_argument_errors_need_no_gpu():
    """The two-phase temporary-storage protocol (dispatch_radix_sort.cuh:846-850) is pure host arithmetic."""
    lib = ctypes.CDLL(LIB)
    vp, sz, u64, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
    lib.b200_lsb_sort.restype = i32
    lib.b200_lsb_sort.argtypes = [vp, ctypes.POINTER(sz), vp, vp, vp, vp, ctypes.POINTER(i32), u64, i32, i32, i32, i32, i32, i32, vp]
    lib.b200_msb_sort.restype = i32
    lib.b200_msb_sort.argtypes = [vp, vp, u64, vp, vp, i32, i32, vp, ctypes.POINTER(sz), vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    prev = 0
    for n in (0, 1, 1000, 1 << 20, 1 << 28, 1 << 32):
        for kt, vb in ((0, 0), (0, 4), (1, 0), (1, 8), (4, 4), (5, 0)):
            b = sz(0)
            assert lib.b200_lsb_sort(None, ctypes.byref(b), None, None, None, None, None, n, kt, vb, 0, 64, 0, 1, None) == 0
            assert b.value >= 256
            b2 = sz(0)
            assert lib.b200_lsb_sort(None, ctypes.byref(b2), None, None, None, None, None, n, kt, vb, 0, 64, 0, 0, None) == 0
            assert b2.value >= b.value            # pointer overloads need the third buffer (dispatch_radix_sort.cuh:1099-1104)
            w = sz(0)
            assert lib.b200_msb_sort(None, None, n, None, None, kt, vb, None, ctypes.byref(w), None, None, None) == 0
            assert w.value >= 256
        prev = n
    # segmented sort: the same two-phase protocol, plus two on-chip work lists sized by the segment count
    lib.b200_segmented_sort.restype = i32
    lib.b200_segmented_sort.argtypes = [vp, ctypes.POINTER(sz), vp, vp, vp, vp, ctypes.POINTER(i32), u64, ctypes.c_uint32, vp, vp, i32,
                                        i32, i32, i32, i32, i32, i32, vp]
    for n, ns in ((0, 0), (1000, 1), (1 << 20, 5000), (1 << 28, 1 << 20)):
        b = sz(0); b1 = sz(0); b0 = sz(0)
        assert lib.b200_segmented_sort(None, ctypes.byref(b), None, None, None, None, None, n, ns, None, None, 4, 0, 4, 0, 32, 0, 1, None) == 0
        assert lib.b200_lsb_sort(None, ctypes.byref(b1), None, None, None, None, None, n, 0, 4, 0, 32, 0, 1, None) == 0
        assert b.value >= b1.value + 2 * 16 * ns
        assert lib.b200_segmented_sort(None, ctypes.byref(b0), None, None, None, None, None, n, ns, None, None, 4, 0, 4, 0, 32, 0, 0, None) == 0
        assert b0.value >= b.value
    b = sz(0)
    assert lib.b200_segmented_sort(None, ctypes.byref(b), None, None, None, None, None, 10, 1, None, None, 2, 0, 0, 0, 32, 0, 1, None) != 0   # bad offset width
    assert lib.b200_segmented_sort(None, ctypes.byref(b), None, None, None, None, None, 1 << 32, 1, None, None, 8, 0, 0, 0, 32, 0, 1, None) != 0   # n >= 2^32
    b = sz(0)
    assert lib.b200_lsb_sort(None, ctypes.byref(b), None, None, None, None, None, 10, 99, 0, 0, 32, 0, 1, None) != 0   # bad key type
    assert lib.b200_lsb_sort(None, ctypes.byref(b), None, None, None, None, None, 10, 0, 3, 0, 32, 0, 1, None) != 0    # bad value width
    assert lib.b200_lsb_sort(None, None, None, None, None, None, None, 10, 0, 0, 0, 32, 0, 1, None) != 0              # no size pointer


def test_product_never_references_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/ (task statement, section 3)."""
    pkg = os.path.join(ROOT, "gpu_sort_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_lib" not in text and "oracle/" not in text.replace("oracle/radix_oracle.c", ""), f
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            text = open(os.path.join(dirpath, f)).read()
            assert "liboracle" not in text


def test_python_host_layer_fails_loudly_without_the_library(tmp_path):
    code = ("import importlib.util,sys,os;"
            f"spec=importlib.util.spec_from_file_location('g', r'{os.path.join(ROOT, 'gpu_sort_b200', '__init__.py')}');"
            "m=importlib.util.module_from_spec(spec);"
            "import builtins; real=os.path.exists;"
            "os.path.exists=lambda p: False if p.endswith('libb200sort.so') else real(p);"
            "\ntry:\n spec.loader.exec_module(m)\n print('LOADED')\nexcept ImportError as e:\n print('IMPORTERROR')\n")
    out = subprocess.check_output(["python", "-c", code], text=True, cwd=str(tmp_path))
    assert "IMPORTERROR" in out
