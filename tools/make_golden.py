#!/usr/bin/env python
"""Generates tests/golden/ref_digests.json ON THE GPU BOX from the UNMODIFIED reference (oracle/_ref/*.so, compiled
from /root/reference by oracle/Makefile): digests of the reference's own outputs on seeded inputs.

  reference-msb : rdxsrt_unstable_sort (msb/src/sort/gpu_radix_sort.h:187-507)
  reference-lsb : cub::DeviceRadixSort call shape of lsb/sort.cu:25-76 (toolkit CUB 2.8.2; the vendored 1.6.4 cannot
                  be assembled for sm_100)
Inputs come from the portable generator shared by oracle/radix_oracle.c and the library, the
families are the reference tests' own (entropy levels, default sizes 200000 keys / 100000 pairs, msb/tests/*.cu).
usage (via gpurun):  python tools/make_golden.py gpurun_out/ref_digests.json
"""
import ctypes, hashlib, json, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_lib  # noqa: E402

KT = {"u32": 0, "u64": 1, "i32": 2, "i64": 3, "f32": 4, "f64": 5}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def dev(a):
    return torch.from_numpy(a.view(np.int32 if a.dtype.itemsize == 4 else np.int64).copy()).cuda()


def P(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_digests.json")
    orc = oracle_lib.load()
    msb = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libref_msb.so")); msb.ref_msb_sort_device.restype = ctypes.c_int
    lsb = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libref_lsb.so")); lsb.ref_lsb_cub_sort.restype = ctypes.c_int
    cases = []
    # ---- reference MSB: keys-only over entropy levels (test_sort_keys.cu:121-149), pairs (test_sort_pairs.cu:223-281), skew
    spec = []
    for kt in ("u32", "u64"):
        for level in (1, 2, 3, 5, 8, 11, 0):
            spec.append((kt, 0, 200000, "entropy", level))
        for vb in (4, 8):
            spec.append((kt, vb, 100000, "uniform", 0))
            spec.append((kt, vb, 100000, "entropy", 3))
        for dist in ("zipf_rank", "zipf_hash", "sorted", "reverse", "constant"):
            spec.append((kt, 0, 1 << 20, dist, 0))
        spec.append((kt, 0, 1258925, "uniform", 0))          # 100000 * 10^(11/10), the geometric sweep (test_sort_keys.cu:179)
    for kt, vb, n, dist, param in spec:
        bits = 32 if kt.endswith("32") else 64
        k = orc.gen_keys(n, bits, seed=0, dist=dist, param=param)
        v = np.arange(n, dtype=np.uint32 if vb == 4 else np.uint64) if vb else None
        k0 = dev(k); k1 = torch.empty_like(k0)
        v0 = dev(v) if vb else None; v1 = torch.empty_like(v0) if vb else None
        ok, ov = ctypes.c_void_p(0), ctypes.c_void_p(0)
        rc = msb.ref_msb_sort_device(P(k0), P(v0), ctypes.c_ulonglong(n), P(k1), P(v1), ctypes.c_int(bits), ctypes.c_int(vb),
                                     ctypes.byref(ok), ctypes.byref(ov))
        torch.cuda.synchronize()
        rk = (k0 if ok.value == k0.data_ptr() else k1).cpu().numpy().view(k.dtype)
        c = {"impl": "reference-msb", "key_type": kt, "value_bytes": vb, "n": n, "seed": 0, "dist": dist, "param": param,
             "rc": rc, "keys_sha256": sha(rk)}
        if vb:
            rv = (v0 if ov.value == v0.data_ptr() else v1).cpu().numpy().view(v.dtype)
            s, x = orc.digest(rk, rv)
            c["pair_digest"] = [s, x]
        assert np.array_equal(rk, np.sort(k)), f"reference MSB mis-sorted {c}"
        cases.append(c)
    # ---- reference LSB (CUB): all key types, keys and pairs, ascending and descending
    for kt in ("u32", "u64", "i32", "i64", "f32", "f64"):
        bits = 32 if kt.endswith("32") else 64
        for vb, n, dist, param, desc in ((0, 200000, "uniform", 0, False), (4, 100000, "uniform", 0, False), (8, 100000, "entropy", 3, True),
                                          (4, 1 << 20, "zipf_hash", 0, False), (0, 100000, "entropy", 5, True)):
            if kt in ("f32", "f64") and dist == "entropy":
                # AND-ed streams are full of +0.0 / -0.0 bit patterns; the toolkit's CUB 2.x treats -0.0 == +0.0 while the
                # reference (vendored CUB 1.6.4 Traits, pure bit order) does not -- keep floats on uniform bit patterns
                dist, param = "uniform", 0
            k = orc.gen_keys(n, bits, seed=1, dist=dist, param=param).view(oracle_lib.NP_OF[kt])
            v = np.arange(n, dtype=np.uint32 if vb == 4 else np.uint64) if vb else None
            k0 = dev(k); k1 = torch.empty_like(k0)
            v0 = dev(v) if vb else None; v1 = torch.empty_like(v0) if vb else None
            tb = ctypes.c_size_t(0); sel = ctypes.c_int(0)
            args = lambda temp: (temp, ctypes.byref(tb), P(k0), P(k1), P(v0), P(v1), ctypes.c_int(n), ctypes.c_int(KT[kt]),
                                 ctypes.c_int(vb), ctypes.c_int(int(desc)), ctypes.c_int(0), ctypes.c_int(bits), ctypes.byref(sel))
            lsb.ref_lsb_cub_sort(*args(None))
            temp = torch.empty(max(tb.value, 1), dtype=torch.uint8, device="cuda")
            rc = lsb.ref_lsb_cub_sort(*args(P(temp)))
            torch.cuda.synchronize()
            rk = (k1 if sel.value else k0).cpu().numpy().view(k.dtype)
            c = {"impl": "reference-lsb", "key_type": kt, "value_bytes": vb, "n": n, "seed": 1, "dist": dist, "param": param,
                 "descending": desc, "rc": rc, "keys_sha256": sha(rk)}
            if vb:
                c["values_sha256"] = sha((v1 if sel.value else v0).cpu().numpy().view(v.dtype))
            cases.append(c)
    meta = {"generated_by": "tools/make_golden.py", "device": torch.cuda.get_device_name(0),
            "reference": "anilshanbhag/gpu-sort compiled unmodified from /root/reference (oracle/Makefile), toolkit CUB 2.8.2 for the LSB arm",
            "cases": cases}
    with open(out_path, "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", out_path, len(cases), "cases")


if __name__ == "__main__":
    main()
