"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI
(gpu_sort_b200 binds include/b200sort.h with ctypes), against
  * the CPU oracle (oracle/radix_oracle.c) on the same seeded inputs -- bit-exact keys; bit-exact values for the stable
    LSB path; identical (key, value) multiset for the unstable MSB path (the reference's own criterion,
    msb/tests/test_sort_pairs.cu:80-109,166-176);
  * tests/golden/ref_digests.json -- digests of the outputs of the UNMODIFIED reference on a B200;
  * size-independent properties at BASELINE.json's full sizes (sortedness, multiset digest, stability of iota values).
The test families are the reference's: entropy levels {1..11,0} (msb/tests/test_sort_keys.cu:121-149), default sizes
200000 keys / 100000 pairs, the geometric size sweep (:179), key/value type matrix (test_sort_pairs.cu:223-281), and
CUB's matrix: descending, bit sub-ranges, pointer (non-overwriting) overloads, n -> ceil(n/32) ... 1, 0
(lsb/cub/test/test_device_radix_sort.cu:956-1080).
"""
import hashlib
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests.oracle_lib import NP_OF  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_digests.json")
KT_ID = {"u32": 0, "u64": 1, "i32": 2, "i64": 3, "f32": 4, "f64": 5}


@pytest.fixture(scope="module")
def gs():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import gpu_sort_b200 as g          # raises if libb200sort.so is missing: there is no fallback path
    torch.cuda.set_device(0)
    return g


def dev(a):
    if a is None:
        return None
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32 if a.dtype.itemsize == 4 else np.int64).copy()).cuda()


def host(t, dtype):
    return t.cpu().numpy().view(dtype)


def raw_keys(orc, n, kt, seed=0, dist="uniform", param=0):
    bits = 32 if kt.endswith("32") else 64
    return orc.gen_keys(n, bits, seed=seed, dist=dist, param=param).view(NP_OF[kt])


def iota(n, vb):
    return np.arange(n, dtype=np.uint32 if vb == 4 else np.uint64) if vb else None


def run_lsb(gs, k, v, kt, descending=False, begin_bit=0, end_bit=None, overwrite=True):
    n = k.size
    k0, k1 = dev(k), torch.empty_like(dev(k))
    v0 = dev(v); v1 = torch.empty_like(v0) if v is not None else None
    fn = {(False, False): gs.DeviceRadixSort.SortKeys, (False, True): gs.DeviceRadixSort.SortKeysDescending,
          (True, False): gs.DeviceRadixSort.SortPairs, (True, True): gs.DeviceRadixSort.SortPairsDescending}[(v is not None, descending)]
    if overwrite:
        dk = gs.DoubleBuffer(k0, k1); dv = gs.DoubleBuffer(v0, v1) if v is not None else None
        args = (dk, dv, n) if v is not None else (dk, n)
        kw = dict(begin_bit=begin_bit, end_bit=end_bit, key_type=KT_ID[kt])
    else:
        args = (k0, v0, n) if v is not None else (k0, n)
        kw = dict(begin_bit=begin_bit, end_bit=end_bit, key_type=KT_ID[kt], d_keys_out=k1)
        if v is not None:
            kw["d_values_out"] = v1
    tb = fn(None, *args, **kw)
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    fn(temp, *args, **kw)
    torch.cuda.synchronize()
    if overwrite:
        rk = host(dk.Current(), k.dtype); rv = host(dv.Current(), v.dtype) if v is not None else None
    else:
        assert np.array_equal(host(k0, k.dtype).view(np.uint8), k.view(np.uint8)), "pointer overload must not touch the input"
        rk = host(k1, k.dtype); rv = host(v1, v.dtype) if v is not None else None
    return rk, rv


def run_msb(gs, k, v, kt, workspace=False):
    n = k.size
    k0, k1 = dev(k), torch.empty_like(dev(k))
    v0 = dev(v); v1 = torch.empty_like(v0) if v is not None else None
    ws = None
    if workspace:
        ws = torch.empty(gs.rdxsrt_workspace_bytes(n, KT_ID[kt], 0 if v is None else v.dtype.itemsize), dtype=torch.uint8, device="cuda")
    r = gs.rdxsrt_unstable_sort(k0, v0, n, k1, v1, workspace=ws, key_type=KT_ID[kt])
    torch.cuda.synchronize()
    if n:   # the reference returns the INPUT buffers for 4/8-byte keys (gpu_radix_sort.h:359-360)
        assert r.sorted_keys.data_ptr() == k0.data_ptr()
    return host(r.sorted_keys, k.dtype), host(r.sorted_values, v.dtype) if v is not None else None


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def pair_multiset(k, v):
    ku = k.view(np.uint32 if k.dtype.itemsize == 4 else np.uint64)
    order = np.lexsort((v, ku))
    return ku[order], v[order]


# ------------------------------------------------------------------------------------------------------------------
# stable LSB path vs oracle: bit-exact keys AND values
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kt", ["u32", "u64", "i32", "i64", "f32", "f64"])
@pytest.mark.parametrize("vb", [0, 4, 8])
def test_lsb_type_matrix(gs, oracle, kt, vb):
    for n in (100000, 6145):
        k = raw_keys(oracle, n, kt, seed=1)
        v = iota(n, vb)
        ek, ev = oracle.lsb_sort(k, v, key_type=kt)
        rk, rv = run_lsb(gs, k, v, kt)
        assert same_bits(rk, ek)
        if vb:
            assert np.array_equal(rv, ev)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 33, 1000, 6144, 6145, 8192, 8193, 24577, 200000, (1 << 20) + 3, 3211264 + 77])
def test_lsb_sizes_pairs(gs, oracle, n):
    k = raw_keys(oracle, n, "u32", seed=2)
    v = iota(n, 4)
    ek, ev = oracle.lsb_sort(k, v, key_type="u32")
    rk, rv = run_lsb(gs, k, v, "u32")
    assert same_bits(rk, ek) and np.array_equal(rv, ev)


@pytest.mark.parametrize("level", [0, 1, 2, 3, 5, 8, 11])
@pytest.mark.parametrize("kt", ["u32", "u64"])
def test_lsb_entropy_levels_stability(gs, oracle, kt, level):
    n = 300000
    k = raw_keys(oracle, n, kt, seed=0, dist="entropy", param=level)
    v = iota(n, 4)
    ek, ev = oracle.lsb_sort(k, v, key_type=kt)
    rk, rv = run_lsb(gs, k, v, kt)
    assert same_bits(rk, ek) and np.array_equal(rv, ev)     # equal keys keep input order


@pytest.mark.parametrize("kt", ["u32", "i32", "f32", "u64", "f64"])
@pytest.mark.parametrize("vb", [0, 4])
def test_lsb_descending(gs, oracle, kt, vb):
    n = 150001
    k = raw_keys(oracle, n, kt, seed=3, dist="entropy", param=2 if kt[0] != "f" else 0) if kt[0] != "f" else raw_keys(oracle, n, kt, seed=3)
    v = iota(n, vb)
    ek, ev = oracle.lsb_sort(k, v, key_type=kt, descending=True)
    rk, rv = run_lsb(gs, k, v, kt, descending=True)
    assert same_bits(rk, ek)
    if vb:
        assert np.array_equal(rv, ev)


@pytest.mark.parametrize("bits", [(0, 8), (4, 20), (15, 17), (1, 31), (24, 32), (3, 3), (0, 13)])
@pytest.mark.parametrize("n", [5000, 250000])
def test_lsb_bit_subranges(gs, oracle, bits, n):
    b, e = bits
    k = raw_keys(oracle, n, "u32", seed=5)
    v = iota(n, 8)
    ek, ev = oracle.lsb_sort(k, v, key_type="u32", begin_bit=b, end_bit=e)
    rk, rv = run_lsb(gs, k, v, "u32", begin_bit=b, end_bit=e)
    assert same_bits(rk, ek) and np.array_equal(rv, ev)


def test_lsb_bit_subrange_u64(gs, oracle):
    k = raw_keys(oracle, 180000, "u64", seed=6)
    v = iota(k.size, 4)
    for b, e in ((0, 64), (31, 33), (8, 40), (60, 64)):
        ek, ev = oracle.lsb_sort(k, v, key_type="u64", begin_bit=b, end_bit=e)
        rk, rv = run_lsb(gs, k, v, "u64", begin_bit=b, end_bit=e)
        assert same_bits(rk, ek) and np.array_equal(rv, ev)


@pytest.mark.parametrize("kt,vb,n", [("u32", 4, 200000), ("u32", 0, 3000), ("u64", 8, 70000), ("f32", 4, 100000), ("u32", 4, 0)])
def test_lsb_pointer_overloads_leave_input_untouched(gs, oracle, kt, vb, n):
    k = raw_keys(oracle, n, kt, seed=7)
    v = iota(n, vb)
    ek, ev = oracle.lsb_sort(k, v, key_type=kt)
    rk, rv = run_lsb(gs, k, v, kt, overwrite=False)
    assert same_bits(rk, ek)
    if vb:
        assert np.array_equal(rv, ev)


@pytest.mark.parametrize("dist", ["zipf_rank", "zipf_hash", "sorted", "reverse", "constant"])
def test_lsb_skewed(gs, oracle, dist):
    n = 400000
    k = raw_keys(oracle, n, "u64", seed=2, dist=dist)
    v = iota(n, 4)
    ek, ev = oracle.lsb_sort(k, v, key_type="u64")
    rk, rv = run_lsb(gs, k, v, "u64")
    assert same_bits(rk, ek) and np.array_equal(rv, ev)


def test_lsb_reference_driver_shape_float_keys_random_values(gs, oracle):
    """lsb/sort.cu:110-152: float keys in (0,1], random uint values, SortPairs then SortKeysDescending (quirk Q1)."""
    n = 1 << 20
    rng = np.random.default_rng(0)
    k = (1.0 - rng.random(n, dtype=np.float32)).astype(np.float32)
    v = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ek, ev = oracle.lsb_sort(k, v, key_type="f32")
    rk, rv = run_lsb(gs, k, v, "f32")
    assert same_bits(rk, ek) and np.array_equal(rv, ev)
    ek, _ = oracle.lsb_sort(k, None, key_type="f32", descending=True)
    rk, _ = run_lsb(gs, k, None, "f32", descending=True)
    assert same_bits(rk, ek)


# ------------------------------------------------------------------------------------------------------------------
# unstable MSB path vs oracle: bit-exact keys, identical (key, value) multiset
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0])        # test_sort_keys.cu:126
@pytest.mark.parametrize("kt", ["u32", "u64", "f64"])                           # Entropy_UINT / _ULONG / _DOUBLE (:154-170)
def test_msb_keys_entropy_levels(gs, oracle, kt, level):
    n = 200000
    k = raw_keys(oracle, n, kt, seed=0, dist="entropy", param=level)
    ek, _ = oracle.msb_sort(k, key_type=kt)
    rk, _ = run_msb(gs, k, None, kt)
    assert same_bits(rk, ek)


@pytest.mark.parametrize("kt,vb", [("u32", 4), ("u32", 8), ("u64", 4), ("u64", 8)])     # test_sort_pairs.cu:223-253
@pytest.mark.parametrize("level", [1, 3, 6, 0])
def test_msb_pairs(gs, oracle, kt, vb, level):
    n = 100000
    k = raw_keys(oracle, n, kt, seed=0, dist="entropy", param=level)
    v = iota(n, vb)
    ek, ev = oracle.msb_sort(k, v, key_type=kt)
    rk, rv = run_msb(gs, k, v, kt)
    assert same_bits(rk, ek)
    # reference fast check (test_sort_pairs.cu:166-176) ...
    assert np.array_equal(k[rv.astype(np.int64)], rk) and int(rv.astype(np.uint64).sum()) == n * (n - 1) // 2
    # ... and the exact multiset against the oracle's output
    a, b = pair_multiset(rk, rv), pair_multiset(ek, ev)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert oracle.digest(rk, rv) == oracle.digest(k, v)


@pytest.mark.parametrize("n", [0, 1, 2, 33, 1000, 6144, 6145, 12289, 100000, 125892, 1258925, (1 << 22) + 5])
@pytest.mark.parametrize("workspace", [False, True])
def test_msb_sizes(gs, oracle, n, workspace):
    k = raw_keys(oracle, n, "u32", seed=0)
    ek, _ = oracle.msb_sort(k, key_type="u32")
    rk, _ = run_msb(gs, k, None, "u32", workspace=workspace)
    assert same_bits(rk, ek)


@pytest.mark.parametrize("kt", ["i32", "i64", "f32", "f64"])
def test_msb_signed_and_float_keys(gs, oracle, kt):
    n = 300000
    k = raw_keys(oracle, n, kt, seed=4)
    v = iota(n, 4)
    ek, ev = oracle.msb_sort(k, v, key_type=kt)
    rk, rv = run_msb(gs, k, v, kt)
    assert same_bits(rk, ek)
    assert same_bits(k[rv.astype(np.int64)], rk)


@pytest.mark.parametrize("dist", ["zipf_rank", "zipf_hash", "sorted", "reverse", "constant"])
@pytest.mark.parametrize("kt", ["u32", "u64"])
def test_msb_skewed(gs, oracle, kt, dist):
    n = (1 << 21) + 17
    k = raw_keys(oracle, n, kt, seed=2, dist=dist)
    v = iota(n, 4)
    ek, _ = oracle.msb_sort(k, None, key_type=kt)
    rk, rv = run_msb(gs, k, v, kt)
    assert same_bits(rk, ek)
    assert np.array_equal(k[rv.astype(np.int64)], rk)
    assert oracle.digest(rk, rv) == oracle.digest(k, v)


def test_host_pointer_wrappers(gs, oracle):
    """rdxsrt_unstable_sort_keys / _pairs (gpu_radix_sort.h:510-587) and the LSB twin."""
    n = 250000
    k = raw_keys(oracle, n, "u32", seed=9)
    v = iota(n, 4)
    ek, _ = oracle.msb_sort(k, key_type="u32")
    assert same_bits(gs.rdxsrt_unstable_sort_keys(k), ek)
    rk, rv = gs.rdxsrt_unstable_sort_pairs(k, v)
    assert same_bits(rk, ek) and np.array_equal(k[rv.astype(np.int64)], rk)
    k64 = raw_keys(oracle, n, "u64", seed=9)
    assert same_bits(gs.rdxsrt_unstable_sort_keys(k64), np.sort(k64))
    ek, ev = oracle.lsb_sort(k, v, key_type="u32", descending=True)
    rk, rv = gs.lsb_sort_host(k, v, descending=True)
    assert same_bits(rk, ek) and np.array_equal(rv, ev)


# ------------------------------------------------------------------------------------------------------------------
# golden: digests of the UNMODIFIED reference's outputs on a B200 (tests/golden/ref_digests.json)
# ------------------------------------------------------------------------------------------------------------------
def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def test_cuda_path_matches_reference_golden(gs, oracle):
    gold = json.load(open(GOLDEN))
    assert len(gold["cases"]) >= 60
    for c in gold["cases"]:
        kt = c["key_type"]
        k = raw_keys(oracle, c["n"], kt, seed=c["seed"], dist=c["dist"], param=c["param"])
        v = iota(c["n"], c["value_bytes"])
        if c["impl"] == "reference-msb":
            rk, rv = run_msb(gs, k, v, kt)
            assert _sha(rk) == c["keys_sha256"], c
            if v is not None:
                assert list(oracle.digest(rk, rv)) == c["pair_digest"], c
        else:
            rk, rv = run_lsb(gs, k, v, kt, descending=c.get("descending", False))
            assert _sha(rk) == c["keys_sha256"], c
            if v is not None:
                assert _sha(rv) == c["values_sha256"], c


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU building blocks on one device
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kt,vb", [("u32", 4), ("u64", 8), ("u32", 0), ("f32", 4)])
@pytest.mark.parametrize("parts", [1, 2, 3, 8])
def test_histogram_and_range_partition(gs, oracle, kt, vb, parts):
    from gpu_sort_b200 import dist as gd
    n = 700001
    bits = 12
    k = raw_keys(oracle, n, kt, seed=11, dist="entropy", param=2 if kt[0] != "f" else 1)
    v = iota(n, vb)
    from tests.test_oracle import twiddle_np
    tw = twiddle_np(k, kt)
    bucket = (tw >> np.array(tw.dtype.itemsize * 8 - bits, dtype=tw.dtype)).astype(np.int64)
    dk, dv = dev(k), dev(v)
    counts = gd.msd_histogram(dk, bits, key_type=KT_ID[kt])
    assert np.array_equal(counts.cpu().numpy().view(np.uint64), np.bincount(bucket, minlength=1 << bits).astype(np.uint64))
    splitters = gd.choose_splitters(counts.cpu().numpy().view(np.uint64), parts)
    assert len(splitters) == parts - 1
    ok, ov, offs = gd.range_partition(dk, dv, bits, splitters, counts, key_type=KT_ID[kt])
    torch.cuda.synchronize()
    dest = np.searchsorted(np.asarray(splitters, dtype=np.int64), bucket, side="right") if parts > 1 else np.zeros(n, dtype=np.int64)
    order = np.argsort(dest, kind="stable")
    assert same_bits(host(ok, k.dtype), k[order])                      # stable G-way split
    if vb:
        assert np.array_equal(host(ov, v.dtype), v[order])
    exp_offs = np.concatenate([[0], np.cumsum(np.bincount(dest, minlength=parts))]).astype(np.uint64)
    assert np.array_equal(offs.cpu().numpy().view(np.uint64), exp_offs)


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties (sortedness, multiset digest, stability)
# ------------------------------------------------------------------------------------------------------------------
def _full_size(gs, path, logn, key_bits, vb, dist="uniform", param=0):
    n = 1 << logn
    kt = gs.KEY_U32 if key_bits == 32 else gs.KEY_U64
    src = torch.empty(n, dtype=torch.int32 if key_bits == 32 else torch.int64, device="cuda")
    gs.generate_keys(src, seed=0, dist=dist, param=param)
    vsrc = gs.iota(torch.empty(n, dtype=torch.int32 if vb == 4 else torch.int64, device="cuda")) if vb else None
    before = gs.check(src, vsrc, key_type=kt)[:2]
    alt = torch.empty_like(src); valt = torch.empty_like(vsrc) if vb else None
    if path == "lsb":
        dk = gs.DoubleBuffer(src, alt); dv = gs.DoubleBuffer(vsrc, valt) if vb else None
        tb = gs.DeviceRadixSort._run(None, dk, dv, n, 0, None, False, None, kt)
        temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
        gs.DeviceRadixSort._run(temp, dk, dv, n, 0, None, False, None, kt)
        rk, rv = dk.Current(), dv.Current() if vb else None
    else:
        r = gs.rdxsrt_unstable_sort(src, vsrc, n, alt, valt, key_type=kt)
        rk, rv = r.sorted_keys, r.sorted_values
    torch.cuda.synchronize()
    s, x, bad, vbad = gs.check(rk, rv, key_type=kt)
    assert bad == 0, "output not sorted"
    assert (s, x) == before, "(key, value) multiset changed"
    if path == "lsb" and vb:
        assert vbad == 0, "stable sort of iota values must keep values ascending inside equal keys"
    # idempotence: sorting the sorted output again changes nothing
    h0 = (s, x)
    del src, alt
    return h0


def test_full_cfg1_2p24_u32_keys_vs_host_sort(gs, oracle):
    """BASELINE config 1: 2^24 uniform u32 keys-only validated against the host sort, both paths."""
    n = 1 << 24
    k = raw_keys(oracle, n, "u32", seed=0)
    exp = np.sort(k)
    assert same_bits(run_msb(gs, k, None, "u32")[0], exp)
    assert same_bits(run_lsb(gs, k, None, "u32")[0], exp)


def test_full_cfg2_2p28_u32_keys_msb(gs):
    _full_size(gs, "msb", 28, 32, 0)


def test_full_cfg3_2p28_u32_pairs_lsb(gs):
    _full_size(gs, "lsb", 28, 32, 4)


@pytest.mark.parametrize("dist,param", [("uniform", 0), ("zipf_rank", 0), ("zipf_hash", 0), ("entropy", 3), ("sorted", 0), ("reverse", 0), ("constant", 0)])
def test_full_cfg4_2p29_u64_skewed_msb(gs, dist, param):
    _full_size(gs, "msb", 29, 64, 0, dist, param)


def test_full_cfg4_2p29_u64_lsb(gs):
    _full_size(gs, "lsb", 29, 64, 0, "zipf_hash")


def test_dist_sorter_single_rank(gs, oracle, tmp_path):
    """gpu_sort_b200.dist.DistSorter with a one-rank process group: histogram, splitter choice, range partition, local sort with
    the key-range bit hint -- the whole multi-GPU code path except the peer-memory stores (bench.py --gpus N checks those)."""
    import torch.distributed as dist
    from gpu_sort_b200 import dist as gd
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        for pairs, n in ((True, (1 << 20) + 77), (False, 3_000_001)):
            k = raw_keys(oracle, n, "u32", seed=13, dist="entropy", param=1)
            v = iota(n, 4) if pairs else None
            sorter = gd.DistSorter(n, torch.int32, torch.int32 if pairs else None, key_type=KT_ID["u32"])
            sk, sv, info = sorter.sort(dev(k), dev(v))
            torch.cuda.synchronize()
            assert info["count"] == n
            ek, ev = oracle.lsb_sort(k, v, key_type="u32")
            assert same_bits(host(sk, k.dtype), ek)
            if pairs:
                assert np.array_equal(host(sv, v.dtype), ev)        # stable end to end
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kt,vb", [("u32", 0), ("u32", 4), ("u64", 0), ("u64", 8), ("i32", 4), ("f32", 0)])
def test_key_range_probe_paths(gs, oracle, kt, vb):
    """n >= 2^22 switches the key-range probe on: leading bits that are equal in all keys are skipped and the result buffer
    follows the number of levels actually run.  Small-range keys, an all-equal array, and full-range keys, through both entry
    points (overwriting and pointer overloads)."""
    n = (1 << 22) + 3
    bits = 32 if kt.endswith("32") else 64
    base = oracle.gen_keys(n, bits, seed=21)
    variants = {
        "small_range": base & np.array((1 << 19) - 1, dtype=base.dtype),
        "high_bits_fixed": (base & np.array((1 << 27) - 1, dtype=base.dtype)) | np.array(0x5 << (bits - 4), dtype=base.dtype),
        "all_equal": np.full(n, 0x1234567, dtype=base.dtype),
        "full_range": base,
    }
    for name, ku in variants.items():
        if kt == "f32" and name != "full_range":
            ku = ku | np.array(0x3F000000, dtype=ku.dtype)       # keep the patterns ordinary positive floats
        k = ku.view(NP_OF[kt])
        v = iota(n, vb)
        ek, ev = oracle.lsb_sort(k, v, key_type=kt)
        for overwrite in (True, False):
            rk, rv = run_lsb(gs, k, v, kt, overwrite=overwrite)
            assert same_bits(rk, ek), (name, overwrite)
            if vb:
                assert np.array_equal(rv, ev), (name, overwrite)
        k0, k1 = dev(k), torch.empty_like(dev(k))
        v0 = dev(v); v1 = torch.empty_like(v0) if vb else None
        r = gs.rdxsrt_unstable_sort(k0, v0, n, k1, v1, key_type=KT_ID[kt])
        torch.cuda.synchronize()
        assert same_bits(host(r.sorted_keys, k.dtype), ek), name
        if vb:
            rv = host(r.sorted_values, v.dtype)
            assert same_bits(k[rv.astype(np.int64)], ek), name


# ------------------------------------------------------------------------------------------------------------------
# segmented sort (cub::DeviceSegmentedRadixSort call shape) vs the oracle: bit-exact keys and values inside every segment.
# Families of lsb/cub/test/test_device_radix_sort.cu:1003-1026 (segment counts n_seg -> ceil(n_seg/32) ... with random
# lengths, single segment), plus segments larger than the on-chip capacity, gaps, empty segments and 64-bit offsets.
# ------------------------------------------------------------------------------------------------------------------
def run_segmented(gs, k, v, kt, begin, end, descending=False, begin_bit=0, end_bit=None, overwrite=True, offset_dtype=np.int32):
    n = k.size
    k0, k1 = dev(k), torch.empty_like(dev(k))
    v0 = dev(v); v1 = torch.empty_like(v0) if v is not None else None
    k1.zero_()
    db = torch.from_numpy(np.ascontiguousarray(begin, dtype=offset_dtype)).cuda()
    de = torch.from_numpy(np.ascontiguousarray(end, dtype=offset_dtype)).cuda()
    S = gs.DeviceSegmentedRadixSort
    fn = {(False, False): S.SortKeys, (False, True): S.SortKeysDescending, (True, False): S.SortPairs, (True, True): S.SortPairsDescending}[(v is not None, descending)]
    ns = len(begin)
    if overwrite:
        dk = gs.DoubleBuffer(k0, k1); dv = gs.DoubleBuffer(v0, v1) if v is not None else None
        args = (dk, dv, n, ns, db, de) if v is not None else (dk, n, ns, db, de)
        kw = dict(begin_bit=begin_bit, end_bit=end_bit, key_type=KT_ID[kt])
    else:
        args = (k0, v0, n, ns, db, de) if v is not None else (k0, n, ns, db, de)
        kw = dict(begin_bit=begin_bit, end_bit=end_bit, key_type=KT_ID[kt], d_keys_out=k1)
        if v is not None:
            kw["d_values_out"] = v1
    tb = fn(None, *args, **kw)
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    fn(temp, *args, **kw)
    torch.cuda.synchronize()
    if overwrite:
        return host(dk.Current(), k.dtype), host(dv.Current(), v.dtype) if v is not None else None
    assert np.array_equal(host(k0, k.dtype).view(np.uint8), k.view(np.uint8)), "pointer overload must not touch the input"
    return host(k1, k.dtype), host(v1, v.dtype) if v is not None else None


def assert_segments_equal(rk, rv, ek, ev, begin, end):
    """Only elements inside a segment are specified (as in CUB)."""
    covered = np.zeros(rk.size, dtype=bool)
    for b, e in zip(begin, end):
        if e > b:
            covered[b:e] = True
    assert same_bits(rk[covered], ek[covered])
    if rv is not None:
        assert np.array_equal(rv[covered], ev[covered])


def cub_segments(n, num_segments, seed):
    rng = np.random.default_rng(seed)
    expected = (n + num_segments - 1) // max(num_segments, 1)
    off = np.zeros(num_segments + 1, dtype=np.int64)
    cur = 0
    for i in range(num_segments):
        off[i] = cur
        cur = min(cur + int(rng.integers(0, 2 * expected + 1)), n)
    off[num_segments] = n
    return off


@pytest.mark.parametrize("kt,vb", [("u32", 0), ("u32", 4), ("u64", 8), ("f32", 4), ("i64", 0), ("f64", 4)])
@pytest.mark.parametrize("descending", [False, True])
def test_segmented_cub_families(gs, oracle, kt, vb, descending):
    n = 300000
    k = raw_keys(oracle, n, kt, seed=4, dist="entropy", param=2)
    v = iota(n, vb)
    num_segments = 5000
    while True:
        off = cub_segments(n, num_segments, seed=num_segments)
        ek, ev = oracle.segmented_sort(k, v, off[:-1], off[1:], key_type=kt, descending=descending)
        rk, rv = run_segmented(gs, k, v, kt, off[:-1], off[1:], descending=descending)
        assert_segments_equal(rk, rv, ek, ev, off[:-1], off[1:])
        if num_segments == 1:
            break
        num_segments = (num_segments + 31) // 32 if num_segments > 32 else 1


@pytest.mark.parametrize("kt,vb", [("u32", 4), ("u64", 0), ("u64", 4)])
@pytest.mark.parametrize("overwrite", [True, False])
def test_segmented_mixed_sizes_gaps_and_empty(gs, oracle, kt, vb, overwrite):
    """Segments from 1 key to far beyond the on-chip capacity in one call, with gaps, empty segments and int64 offsets."""
    n = 1 << 21
    k = raw_keys(oracle, n, kt, seed=9)
    v = iota(n, vb)
    begin = np.array([0, 5, 700, 700, 10000, 50000, 60000, 400000, 1500000, 2000000, 2097151], dtype=np.int64)
    end = np.array([1, 600, 700, 650, 14000, 59000, 390000, 1400000, 1500007, 2097100, 2097152], dtype=np.int64)
    ek, ev = oracle.segmented_sort(k, v, begin, end, key_type=kt)
    rk, rv = run_segmented(gs, k, v, kt, begin, end, overwrite=overwrite, offset_dtype=np.int64)
    assert_segments_equal(rk, rv, ek, ev, begin, end)


def test_segmented_bit_subrange_and_duplicates(gs, oracle):
    n = 200000
    k = raw_keys(oracle, n, "u32", seed=3, dist="entropy", param=4)     # heavy duplicates
    v = iota(n, 4)
    off = cub_segments(n, 40, seed=7)
    for bb, eb in ((0, 32), (8, 24), (15, 17), (5, 5)):
        ek, ev = oracle.segmented_sort(k, v, off[:-1], off[1:], key_type="u32", begin_bit=bb, end_bit=eb)
        rk, rv = run_segmented(gs, k, v, "u32", off[:-1], off[1:], begin_bit=bb, end_bit=eb, overwrite=False)
        assert_segments_equal(rk, rv, ek, ev, off[:-1], off[1:])


def test_segmented_many_tiny_segments(gs, oracle):
    """2^17 segments of 0..16 keys: one launch, no per-segment host work."""
    n = 1 << 20
    k = raw_keys(oracle, n, "u32", seed=11)
    v = iota(n, 4)
    off = cub_segments(n, 1 << 17, seed=5)
    ek, ev = oracle.segmented_sort(k, v, off[:-1], off[1:], key_type="u32")
    rk, rv = run_segmented(gs, k, v, "u32", off[:-1], off[1:])
    assert_segments_equal(rk, rv, ek, ev, off[:-1], off[1:])
    rk, rv = run_segmented(gs, k, None, "u32", np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))      # no segments: a no-op


# ------------------------------------------------------------------------------------------------------------------
# The full-size tests above rest on the product's own generator and checker (b200_util_generate_keys / b200_util_check):
# pin both to the oracle, and show that the checker does flag a broken result.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dist,param", [("uniform", 0), ("entropy", 3), ("entropy", 0), ("zipf_rank", 0), ("zipf_hash", 0), ("sorted", 0), ("reverse", 0), ("constant", 0)])
@pytest.mark.parametrize("bits", [32, 64])
def test_device_generator_equals_oracle_generator(gs, oracle, dist, param, bits):
    n = 300007
    for seed, start, total in ((0, 0, None), (2, 12345, 10 * n)):
        t = torch.empty(n, dtype=torch.int32 if bits == 32 else torch.int64, device="cuda")
        gs.generate_keys(t, seed=seed, dist=dist, param=param, start=start, total=total)
        exp = oracle.gen_keys(n, bits, seed=seed, dist=dist, param=param, start=start, total=total)
        assert np.array_equal(host(t, exp.dtype), exp)


@pytest.mark.parametrize("kt,vb", [("u32", 0), ("u32", 4), ("u64", 8), ("f32", 4), ("i64", 0)])
def test_device_checker_equals_oracle_and_flags_corruption(gs, oracle, kt, vb):
    n = 200003
    k = raw_keys(oracle, n, kt, seed=5, dist="entropy", param=2)
    k[n // 2:n // 2 + n // 4] = k[:n // 4]                                  # plenty of duplicates, whatever the key width
    v = iota(n, vb)
    ek, ev = oracle.lsb_sort(k, v, key_type=kt)
    dk, dv = dev(ek), dev(ev)
    s, x, bad, vbad = gs.check(dk, dv, key_type=KT_ID[kt])
    assert (s, x) == oracle.digest(ek, ev) == oracle.digest(k, v)
    assert bad == 0 == oracle.count_unsorted(ek, kt) and vbad == 0
    # two different keys swapped: same multiset of keys, order broken
    order = np.argsort(ek.view(np.uint32 if ek.dtype.itemsize == 4 else np.uint64), kind="stable")
    i, j = int(order[0]), int(order[-1])
    sw = ek.copy(); sw[[i, j]] = sw[[j, i]]
    swv = ev.copy() if ev is not None else None
    if swv is not None:
        swv[[i, j]] = swv[[j, i]]
    s2, x2, bad2, _ = gs.check(dev(sw), dev(swv), key_type=KT_ID[kt])
    assert bad2 > 0 and bad2 == oracle.count_unsorted(sw, kt)
    assert (s2, x2) == (s, x)                                              # (a swap keeps the multiset: only the order check can see it)
    # one key overwritten: sortedness may survive, the multiset digest must not
    q = int(np.nonzero(ek[1:] != ek[:-1])[0][0]) + 1                       # first position whose key differs from its predecessor
    ck = ek.copy(); ck[q] = ck[q - 1]
    s3, x3, _, _ = gs.check(dev(ck), dv, key_type=KT_ID[kt])
    assert (s3, x3) != (s, x) and (s3, x3) == oracle.digest(ck, ev)
    if vb:
        # the values of two EQUAL keys reversed: keys still sorted, multiset intact, stability broken
        eq = np.nonzero(ek[1:] == ek[:-1])[0]
        assert eq.size, "the test input must contain duplicate keys"
        p = int(eq[0])
        rv = ev.copy(); rv[[p, p + 1]] = rv[[p + 1, p]]
        s4, x4, bad4, vbad4 = gs.check(dk, dev(rv), key_type=KT_ID[kt])
        assert bad4 == 0 and (s4, x4) == (s, x) and vbad4 == 1


def test_full_cfg2_2p28_u32_keys_bit_for_bit_vs_host_sort(gs, oracle):
    """BASELINE config 2 at its full size, compared element by element with a host sort of the same keys (once: ~1 GiB)."""
    n = 1 << 28
    k = oracle.gen_keys(n, 32, seed=0, dist="uniform")
    got = run_msb(gs, k, None, "u32")[0]
    k.sort()
    assert np.array_equal(got, k)


# ------------------------------------------------------------------------------------------------------------------
# The onesweep LSD engine (one up-front histogram read, one decoupled-look-back scatter per digit): selected by
# B200SORT_LSB_ENGINE=onesweep (read once per process, hence the subprocess) and the only engine for n >= 2^32.
# ------------------------------------------------------------------------------------------------------------------
_ONESWEEP_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %(root)r)
import gpu_sort_b200 as gs
from tests import oracle_lib
from tests.test_gpu_parity import run_lsb, raw_keys, iota, same_bits, NP_OF
orc = oracle_lib.load()
for kt, vb, n in (("u32", 0, 200000), ("u32", 4, 250007), ("u64", 8, 70001), ("f32", 4, 100000), ("i64", 0, 33), ("u32", 4, (1 << 22) + 5), ("f64", 4, 1 << 20)):
    for desc in (False, True):
        k = raw_keys(orc, n, kt, seed=3, dist="entropy", param=2); v = iota(n, vb)
        rk, rv = run_lsb(gs, k, v, kt, descending=desc)
        ek, ev = orc.lsb_sort(k, v, key_type=kt, descending=desc)
        assert same_bits(rk, ek) and (rv is None or np.array_equal(rv, ev)), (kt, vb, n, desc)
k = raw_keys(orc, 300000, "u32", seed=1); v = iota(300000, 4)
for bb, eb in ((4, 20), (0, 8), (15, 17)):
    rk, rv = run_lsb(gs, k, v, "u32", begin_bit=bb, end_bit=eb)
    ek, ev = orc.lsb_sort(k, v, key_type="u32", begin_bit=bb, end_bit=eb)
    assert same_bits(rk, ek) and np.array_equal(rv, ev), (bb, eb)
rk, rv = run_lsb(gs, k, v, "u32", overwrite=False)
ek, ev = orc.lsb_sort(k, v, key_type="u32")
assert same_bits(rk, ek) and np.array_equal(rv, ev)
print("ONESWEEP_OK")
"""


def test_onesweep_engine_matrix_in_subprocess():
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200SORT_LSB_ENGINE="onesweep")
    r = subprocess.run([sys.executable, "-c", _ONESWEEP_SCRIPT % {"root": root}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ONESWEEP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_more_than_2p32_keys(gs):
    """n = 2^32 + 5 uint32 keys (beyond the 32-bit tile offsets of the MSD engine: the onesweep engine in < 2^30-key portions):
    sortedness and multiset digest on the device."""
    free, _ = torch.cuda.mem_get_info()
    n = (1 << 32) + 5
    if free < 3 * 4 * n + (2 << 30):
        pytest.skip("needs ~52 GB of free device memory")
    src = torch.empty(n, dtype=torch.int32, device="cuda")
    gs.generate_keys(src, seed=1, dist="uniform")
    before = gs.check(src, None, key_type=gs.KEY_U32)[:2]
    alt = torch.empty_like(src)
    dk = gs.DoubleBuffer(src, alt)
    tb = gs.DeviceRadixSort._run(None, dk, None, n, 0, None, False, None, gs.KEY_U32)
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    gs.DeviceRadixSort._run(temp, dk, None, n, 0, None, False, None, gs.KEY_U32)
    torch.cuda.synchronize()
    s, x, bad, _ = gs.check(dk.Current(), None, key_type=gs.KEY_U32)
    assert bad == 0 and (s, x) == before
    del src, alt, temp
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------------------------
# the reference's OWN harnesses, compiled unmodified against libb200sort.so through the header
# shims (tools/build_ref_on_b200.sh, run by __graft_entry__.build() where /root/reference exists; the binaries travel
# with the snapshot).  msb/tests/test_sort_keys.cu:154-195, test_sort_pairs.cu:223-281, lsb/sort.cu, msb/src/test.cu.
# ------------------------------------------------------------------------------------------------------------------
def _ref_binary(name):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = os.path.join(root, "oracle", "_ref", name)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (needs /root/reference at build time)")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(root, "gpu_sort_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return p, env, root


def test_reference_gtests_pass_on_this_library(gs):
    import subprocess
    p, env, root = _ref_binary("msb_gtests_on_b200sort")
    r = subprocess.run([p, "--gtest_filter=Sort_Keys.Entropy_*:Sort_Pairs.*", "-k", "200000", "-p", "100000"], env=env, cwd=root, capture_output=True, text=True, timeout=900)
    tail = r.stdout[-3000:]
    assert r.returncode == 0 and "[  PASSED  ]" in r.stdout and "[  FAILED  ]" not in r.stdout, tail + r.stderr[-1000:]


def test_reference_lsb_driver_runs_on_this_library(gs):
    import subprocess
    p, env, root = _ref_binary("lsb_sort_on_b200sort")
    r = subprocess.run([p, "--n=16777216", "--t=2"], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
    assert "time_sort" in r.stdout or "time" in r.stdout.lower(), r.stdout[-2000:]


def test_reference_msb_driver_runs_on_this_library(gs):
    import subprocess
    p, env, root = _ref_binary("msb_test_on_b200sort")
    r = subprocess.run([p], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


# ------------------------------------------------------------------------------------------------------------------
# Keys-only sorts on a bit sub-range are STABLE (keys that tie on the window differ elsewhere and keep their input order,
# cub::DeviceRadixSort::SortKeys); device-side errors are reported; the host-wait switch.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits", [(0, 8), (4, 20), (15, 17), (8, 32), (0, 16), (1, 31)])
@pytest.mark.parametrize("n", [5000, 250000, (1 << 21) + 11])
def test_lsb_keys_only_bit_subranges_are_stable(gs, oracle, bits, n):
    k = raw_keys(oracle, n, "u32", seed=4)
    rk, _ = run_lsb(gs, k, None, "u32", begin_bit=bits[0], end_bit=bits[1])
    ek, _ = oracle.lsb_sort(k, None, key_type="u32", begin_bit=bits[0], end_bit=bits[1])
    assert same_bits(rk, ek)


def test_segmented_keys_only_bit_subrange_is_stable(gs, oracle):
    n = 400000
    k = raw_keys(oracle, n, "u64", seed=6)
    cuts = np.sort(np.random.default_rng(1).integers(0, n, size=40))
    begin = np.concatenate([[0], cuts]).astype(np.int64); end = np.concatenate([cuts, [n]]).astype(np.int64)
    k0 = dev(k); k1 = torch.empty_like(k0)
    dk = gs.DoubleBuffer(k0, k1)
    b, e = torch.from_numpy(begin).cuda(), torch.from_numpy(end).cuda()
    tb = gs.DeviceSegmentedRadixSort.SortKeys(None, dk, n, begin.size, b, e, begin_bit=20, end_bit=44, key_type=KT_ID["u64"])
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    gs.DeviceSegmentedRadixSort.SortKeys(temp, dk, n, begin.size, b, e, begin_bit=20, end_bit=44, key_type=KT_ID["u64"])
    torch.cuda.synchronize()
    ek, _ = oracle.segmented_sort(k, None, begin, end, key_type="u64", begin_bit=20, end_bit=44)
    assert same_bits(host(dk.Current(), k.dtype), ek)
    assert gs.sort_status(temp) == 0


def test_device_side_errors_are_reported(gs, oracle):
    """Segment offsets outside [0, n]: the segment is dropped and b200_sort_status says so (the call itself returns cudaSuccess:
    the offsets live on the device)."""
    n = 100000
    k = raw_keys(oracle, n, "u32", seed=2)
    begin = torch.tensor([0, 50000, 90000], dtype=torch.int64, device="cuda")
    end = torch.tensor([50000, 90000, n + 5], dtype=torch.int64, device="cuda")          # the last segment runs past the end
    k0 = dev(k); k1 = torch.empty_like(k0)
    dk = gs.DoubleBuffer(k0, k1)
    tb = gs.DeviceSegmentedRadixSort.SortKeys(None, dk, n, 3, begin, end, key_type=KT_ID["u32"])
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    gs.DeviceSegmentedRadixSort.SortKeys(temp, dk, n, 3, begin, end, key_type=KT_ID["u32"])
    assert gs.sort_status(temp) & 1
    got = host(dk.Current(), k.dtype)
    assert np.array_equal(got[:50000], np.sort(k[:50000])) and np.array_equal(got[50000:90000], np.sort(k[50000:90000]))      # the valid segments are sorted
    end[2] = n
    dk = gs.DoubleBuffer(dev(k), k1)
    gs.DeviceSegmentedRadixSort.SortKeys(temp, dk, n, 3, begin, end, key_type=KT_ID["u32"])
    assert gs.sort_status(temp) == 0


def test_key_range_probe_switch(gs, oracle):
    """b200_set_key_range_probe(0): the call never waits on the host; results are identical (small-range keys just take more sweeps)."""
    n = (1 << 22) + 77
    k = (raw_keys(oracle, n, "u64", seed=9) & np.uint64(0xFFFFF)).astype(np.uint64)       # keys < 2^20 in 64-bit words
    exp = np.sort(k)
    def msb():
        k0 = dev(k); k1 = torch.empty_like(k0)
        r = gs.rdxsrt_unstable_sort(k0, None, n, k1, None, key_type=KT_ID["u64"])
        torch.cuda.synchronize()
        return host(r.sorted_keys, k.dtype), r.sorted_keys.data_ptr() == k0.data_ptr()
    old = gs.set_key_range_probe(False)
    try:
        got, in_input = msb()
        assert same_bits(got, exp) and in_input          # no probe: all 8 levels are planned, the result lands in the input buffer like the reference's
        assert same_bits(run_lsb(gs, k, None, "u64")[0], exp)
    finally:
        gs.set_key_range_probe(old)
    assert same_bits(msb()[0], exp)


@pytest.mark.parametrize("G,xbits,pairs", [(2, 1, True), (2, 4, True), (4, 2, True), (4, 5, True), (4, 6, True), (2, 8, True), (3, 4, False), (8, 3, True)])
def test_exchange_as_level0_emulated_ranks(gs, oracle, G, xbits, pairs):
    """The multi-GPU exchange (b200_exchange_hist -> count matrix -> b200_exchange_scatter -> b200_segmented_sort) with G ranks
    EMULATED on one GPU: every rank's scatter runs as its own launch over its own slice and writes into G receive buffers that all
    live on this device (the kernels never wait on one another, so this is safe on one GPU).  The concatenated result must equal
    ONE stable sort of the concatenated input, bit for bit (SURVEY.md section 8e "Validation").  xbits <= 5 takes the
    destination-aligned write-out, larger xbits the per-position one."""
    import ctypes
    kt = KT_ID["u32"]; vb = 4 if pairs else 0
    sizes = [150_000 + 7777 * r for r in range(G)]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    total = int(offs[-1])
    keys = raw_keys(oracle, total, "u32", seed=11)
    keys[: total // 3] &= np.uint32(0x3FFFFFFF)                    # some skew on the exchange digit
    vals = np.arange(total, dtype=np.uint32) if pairs else None
    cap = int(max(sizes) * 2.5) + 4096
    P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
    recv_k = [torch.zeros(cap, dtype=torch.int32, device="cuda") for _ in range(G)]
    recv_v = [torch.zeros(cap, dtype=torch.int32, device="cuda") for _ in range(G)] if pairs else None
    dst_k = torch.tensor([t.data_ptr() for t in recv_k], dtype=torch.int64, device="cuda")
    dst_v = torch.tensor([t.data_ptr() for t in recv_v] if pairs else [0] * G, dtype=torch.int64, device="cuda")
    dk = [dev(keys[offs[r]:offs[r + 1]]) for r in range(G)]
    dv = [dev(vals[offs[r]:offs[r + 1]]) for r in range(G)] if pairs else [None] * G
    matrix = torch.zeros(G * 256, dtype=torch.int64, device="cuda")
    temps = []
    for r in range(G):
        nb = ctypes.c_size_t(0)
        gs._check(gs.lib.b200_exchange_hist(None, ctypes.byref(nb), None, sizes[r], kt, vb, xbits, None, None), "size")
        t = torch.empty(nb.value, dtype=torch.uint8, device="cuda"); temps.append((t, nb))
        gs._check(gs.lib.b200_exchange_hist(P(t), ctypes.byref(nb), P(dk[r]), sizes[r], kt, vb, xbits, P(matrix[r * 256:(r + 1) * 256]), None), "hist")
    seg_b = [torch.zeros(256, dtype=torch.int64, device="cuda") for _ in range(G)]
    seg_e = [torch.zeros(256, dtype=torch.int64, device="cuda") for _ in range(G)]
    info = [torch.zeros(8, dtype=torch.int64, device="cuda") for _ in range(G)]
    for r in range(G):
        t, nb = temps[r]
        gs._check(gs.lib.b200_exchange_scatter(P(t), ctypes.byref(nb), P(dk[r]), P(dv[r]), sizes[r], kt, vb, xbits, P(matrix), G, r, cap,
                                               P(dst_k), P(dst_v), P(seg_b[r]), P(seg_e[r]), P(info[r]), None), "scatter")
    torch.cuda.synchronize()
    out_k, out_v = [], []
    for r in range(G):
        h = info[r].cpu().numpy()
        assert h[1] == 0, "the plan must fit the receive capacity"
        n_recv = int(h[0])
        kb = gs.DoubleBuffer(recv_k[r], torch.empty_like(recv_k[r]))
        vbuf = gs.DoubleBuffer(recv_v[r], torch.empty_like(recv_v[r])) if pairs else None
        tb = gs.DeviceSegmentedRadixSort._run(None, kb, vbuf, cap, 256, seg_b[r], seg_e[r], 0, 32 - xbits, False, None, kt, ties_are_equal=True)
        temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
        gs.DeviceSegmentedRadixSort._run(temp, kb, vbuf, cap, 256, seg_b[r], seg_e[r], 0, 32 - xbits, False, None, kt, ties_are_equal=True)
        torch.cuda.synchronize()
        assert gs.sort_status(temp) == 0
        out_k.append(host(kb.Current(), np.uint32)[:n_recv])
        if pairs:
            out_v.append(host(vbuf.Current(), np.uint32)[:n_recv])
    got_k = np.concatenate(out_k)
    assert got_k.size == total
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(got_k, keys[order]), "keys: the ranks' results, concatenated, are the sorted global array"
    if pairs:
        assert np.array_equal(np.concatenate(out_v), vals[order]), "values: ONE stable sort of the concatenated input"
