"""CPU tests: pin the oracle (oracle/radix_oracle.c, the CPU restatement of the reference's two sorts).

The reference stores no golden vectors; its own tests are generative and compare against an independent sorter
(msb/tests/test_sort_keys.cu:47-80 -> cub::DeviceRadixSort + memcmp; lsb/cub/test/test_device_radix_sort.cu:554-696 ->
std::stable_sort).  The oracle is pinned the same two ways:
  * against numpy's sort / stable argsort on the reference's own test families (entropy levels, sizes, types,
    descending, bit sub-ranges), and
  * against tests/golden/ref_digests.json: digests of the outputs of the UNMODIFIED reference (oracle/_ref, compiled
    from /root/reference) run on a B200 on the same seeded inputs (tools/make_golden.py generated the file).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.oracle_lib import NP_OF

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_digests.json")


def twiddle_np(a, key_type):
    """numpy statement of cub::Traits<T>::TwiddleIn (lsb/cub/cub/util_type.cuh:966-974,1009-1017,1079-1089)."""
    if key_type in ("u32", "u64"):
        return a.copy()
    bits = a.dtype.itemsize * 8
    u = a.view(np.uint32 if bits == 32 else np.uint64).copy()
    top = np.array(1 << (bits - 1), dtype=u.dtype)
    if key_type in ("i32", "i64"):
        return u ^ top
    neg = (u & top) != 0
    return np.where(neg, ~u, u ^ top)


def raw_keys(orc, n, key_type, seed, dist="uniform", param=0):
    bits = 32 if key_type.endswith("32") else 64
    return orc.gen_keys(n, bits, seed=seed, dist=dist, param=param).view(NP_OF[key_type])


@pytest.mark.parametrize("key_type", ["u32", "u64", "i32", "i64", "f32", "f64"])
def test_twiddle_roundtrip_and_order(oracle, key_type):
    k = raw_keys(oracle, 5000, key_type, seed=3)
    u = k.view(np.uint32 if k.dtype.itemsize == 4 else np.uint64)
    kt = {"u32": 0, "u64": 1, "i32": 2, "i64": 3, "f32": 4, "f64": 5}[key_type]
    tw = np.array([oracle.lib.oracle_twiddle_in(int(x), kt) for x in u[:500]], dtype=np.uint64)
    back = np.array([oracle.lib.oracle_twiddle_out(int(x), kt) for x in tw], dtype=np.uint64)
    assert np.array_equal(back, u[:500].astype(np.uint64))
    assert np.array_equal(tw, twiddle_np(k[:500], key_type).astype(np.uint64))
    if key_type in ("i32", "i64"):   # transformed unsigned order == native signed order
        assert np.array_equal(np.argsort(tw, kind="stable"), np.argsort(k[:500], kind="stable"))


@pytest.mark.parametrize("key_type", ["u32", "u64", "i32", "i64", "f32", "f64"])
@pytest.mark.parametrize("n", [0, 1, 2, 31, 1000, 20000])
def test_lsb_oracle_matches_stable_sort(oracle, key_type, n):
    k = raw_keys(oracle, n, key_type, seed=1)
    v = np.arange(n, dtype=np.uint32)
    order = np.argsort(twiddle_np(k, key_type), kind="stable")
    ko, vo = oracle.lsb_sort(k, v, key_type=key_type)
    assert np.array_equal(ko.view(np.uint8), k[order].view(np.uint8))     # bitwise (NaN != NaN, test_sort_keys.cu:69-70)
    assert np.array_equal(vo, v[order])
    # descending: reverse, stable sort, reverse (test_device_radix_sort.cu:672-676)
    rk, rv = np.ascontiguousarray(k[::-1]), np.ascontiguousarray(v[::-1])
    o2 = np.argsort(twiddle_np(rk, key_type), kind="stable")
    kd, vd = oracle.lsb_sort(k, v, key_type=key_type, descending=True)
    assert np.array_equal(kd.view(np.uint8), np.ascontiguousarray(rk[o2][::-1]).view(np.uint8))
    assert np.array_equal(vd, rv[o2][::-1])


@pytest.mark.parametrize("bits", [(0, 8), (4, 20), (15, 17), (1, 31), (24, 32)])
def test_lsb_oracle_bit_subrange(oracle, bits):
    b, e = bits
    k = raw_keys(oracle, 30000, "u32", seed=5)
    v = np.arange(k.size, dtype=np.uint64)
    proj = (k >> np.uint32(b)) & np.uint32((1 << (e - b)) - 1)
    order = np.argsort(proj, kind="stable")
    ko, vo = oracle.lsb_sort(k, v, key_type="u32", begin_bit=b, end_bit=e)
    assert np.array_equal(ko, k[order]) and np.array_equal(vo, v[order])


@pytest.mark.parametrize("level", [0, 1, 2, 3, 5, 8, 11])      # entropy levels of msb/tests/test_sort_keys.cu:126
@pytest.mark.parametrize("key_bits", [32, 64])
def test_msb_oracle_entropy_levels(oracle, level, key_bits):
    n = 200000   # the reference's default key count (msb/tests/main.cu, -k default)
    kt = "u32" if key_bits == 32 else "u64"
    k = oracle.gen_keys(n, key_bits, seed=0, dist="entropy", param=level)
    v = np.arange(n, dtype=np.uint32)
    ko, vo = oracle.msb_sort(k, v, key_type=kt)
    assert np.array_equal(ko, np.sort(k))
    # the reference's fast value check (test_sort_pairs.cu:166-176): map[v] == key, v < n, sum(v) = n(n-1)/2
    assert np.array_equal(k[vo], ko) and int(vo.astype(np.uint64).sum()) == n * (n - 1) // 2
    assert oracle.digest(k, v) == oracle.digest(ko, vo)


@pytest.mark.parametrize("dist,param", [("zipf_rank", 0), ("zipf_hash", 0), ("sorted", 0), ("reverse", 0), ("constant", 0)])
def test_msb_oracle_skew(oracle, dist, param):
    k = oracle.gen_keys(150000, 64, seed=2, dist=dist, param=param)
    ko, _ = oracle.msb_sort(k, key_type="u64")
    assert np.array_equal(ko, np.sort(k))
    assert oracle.count_unsorted(ko, "u64") == 0


def test_size_sweep_geometric(oracle):
    # msb/tests/test_sort_keys.cu:179: sizes 100000 * 10^(k/10)
    for kk in range(0, 8, 3):
        n = int(100000 * 10 ** (kk / 10))
        k = oracle.gen_keys(n, 32, seed=0)
        assert np.array_equal(oracle.msb_sort(k)[0], np.sort(k))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden digests not generated yet (tools/make_golden.py on the GPU box)")
def test_oracle_matches_reference_golden(oracle):
    """Outputs of the unmodified reference on a B200 (digests committed under tests/golden/)."""
    gold = json.load(open(GOLDEN))
    assert gold["cases"], "empty golden file"
    for c in gold["cases"]:
        kt = c["key_type"]; bits = 32 if kt.endswith("32") else 64
        k = oracle.gen_keys(c["n"], bits, seed=c["seed"], dist=c["dist"], param=c["param"]).view(NP_OF[kt])
        v = np.arange(c["n"], dtype=np.uint32 if c["value_bytes"] == 4 else np.uint64) if c["value_bytes"] else None
        if c["impl"] == "reference-msb":
            ko, vo = oracle.msb_sort(k, v, key_type=kt)
            assert _sha(ko) == c["keys_sha256"], c
            if v is not None:      # unstable: only the (key, value) multiset is defined
                s, x = oracle.digest(ko, vo)
                assert [s, x] == c["pair_digest"], c
        else:
            ko, vo = oracle.lsb_sort(k, v, key_type=kt, descending=c.get("descending", False))
            assert _sha(ko) == c["keys_sha256"], c
            if v is not None:
                assert _sha(vo) == c["values_sha256"], c


# ------------------------------------------------------------------------------------------------------------------
# segmented sort oracle (cub::DeviceSegmentedRadixSort; CPU solution of lsb/cub/test/test_device_radix_sort.cu:669-676)
# ------------------------------------------------------------------------------------------------------------------
def cub_segments(n, num_segments, seed):
    """Segment offsets the way the reference test draws them (InitializeSegments, lsb/cub/test/test_util.h:1475-1494):
    lengths uniform in [0, 2 * expected], clipped at n; the last segment takes the remainder."""
    rng = np.random.default_rng(seed)
    expected = (n + num_segments - 1) // max(num_segments, 1)
    off = np.zeros(num_segments + 1, dtype=np.int64)
    cur = 0
    for i in range(num_segments):
        off[i] = cur
        cur = min(cur + int(rng.integers(0, 2 * expected + 1)), n)
    off[num_segments] = n
    return off


@pytest.mark.parametrize("key_type", ["u32", "i64", "f32"])
@pytest.mark.parametrize("descending", [False, True])
def test_segmented_oracle_matches_per_segment_stable_sort(oracle, key_type, descending):
    n, ns = 20000, 37
    k = raw_keys(oracle, n, key_type, seed=5, dist="entropy", param=3)
    v = np.arange(n, dtype=np.uint32)
    off = cub_segments(n, ns, seed=1)
    ko, vo = oracle.segmented_sort(k, v, off[:-1], off[1:], key_type=key_type, descending=descending)
    tw = twiddle_np(k, key_type)
    for i in range(ns):
        b, e = int(off[i]), int(off[i + 1])
        t = tw[b:e]
        order = np.argsort(~t if descending else t, kind="stable")
        assert np.array_equal(ko[b:e].view(np.uint8), k[b:e][order].view(np.uint8))
        assert np.array_equal(vo[b:e], v[b:e][order])


def test_segmented_oracle_gaps_and_empty_segments(oracle):
    n = 5000
    k = raw_keys(oracle, n, "u32", seed=2)
    begin = np.array([10, 100, 100, 4000, 300], dtype=np.int64)
    end = np.array([60, 100, 90, 5000, 1000], dtype=np.int64)        # second and third are empty (end <= begin)
    ko, _ = oracle.segmented_sort(k, None, begin, end, key_type="u32")
    exp = k.copy()
    for b, e in zip(begin, end):
        if e > b:
            exp[b:e] = np.sort(k[b:e])
    assert np.array_equal(ko, exp)
