#!/usr/bin/env python
"""Fast GPU sanity sweep used while developing kernels (the real parity suite is tests/)."""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, ".")
import gpu_sort_b200 as gs

torch.cuda.init()
dev = "cuda"
fails = 0


def np_keys(n, bits, seed, dist):
    rng = np.random.default_rng(seed)
    dt = np.uint32 if bits == 32 else np.uint64
    if dist == "uniform":
        return rng.integers(0, 2**bits, size=n, dtype=dt)
    if dist == "lowent":
        a = rng.integers(0, 2**bits, size=n, dtype=dt)
        for _ in range(3):
            a &= rng.integers(0, 2**bits, size=n, dtype=dt)
        return a
    if dist == "const":
        return np.full(n, 12345, dtype=dt)
    if dist == "few":
        return rng.integers(0, 7, size=n, dtype=dt) << (bits - 8) | rng.integers(0, 3, size=n, dtype=dt)
    raise ValueError(dist)


def to_dev(a):
    return torch.from_numpy(a.view(np.int32 if a.dtype.itemsize == 4 else np.int64)).to(dev)


def run_case(path, n, bits, vb, dist, seed=0):
    global fails
    keys = np_keys(n, bits, seed, dist)
    kt = gs.KEY_U32 if bits == 32 else gs.KEY_U64
    vals = None
    if vb:
        vals = np.arange(n, dtype=np.uint32 if vb == 4 else np.uint64)
    order = np.argsort(keys, kind="stable")
    exp_k = keys[order]
    k0 = to_dev(keys); k1 = torch.empty_like(k0)
    v0 = to_dev(vals) if vb else None
    v1 = torch.empty_like(v0) if vb else None
    if path == "lsb":
        dk = gs.DoubleBuffer(k0, k1); dv = gs.DoubleBuffer(v0, v1) if vb else None
        tb = gs.DeviceRadixSort._run(None, dk, dv, n, 0, None, False, None, kt)
        temp = torch.empty(tb, dtype=torch.uint8, device=dev)
        gs.DeviceRadixSort._run(temp, dk, dv, n, 0, None, False, None, kt)
        torch.cuda.synchronize()
        rk = dk.Current().cpu().numpy().view(keys.dtype)
        rv = dv.Current().cpu().numpy().view(vals.dtype) if vb else None
    else:
        r = gs.rdxsrt_unstable_sort(k0, v0, n, k1, v1, key_type=kt)
        torch.cuda.synchronize()
        rk = r.sorted_keys.cpu().numpy().view(keys.dtype)
        rv = r.sorted_values.cpu().numpy().view(vals.dtype) if vb else None
    ok = np.array_equal(rk, exp_k)
    okv = True
    if vb and ok:
        if path == "lsb":
            okv = np.array_equal(rv, vals[order])
        else:   # unstable: (key,value) multiset, i.e. keys[rv] == rk and rv is a permutation
            okv = np.array_equal(keys[rv.astype(np.int64)], rk) and np.array_equal(np.sort(rv), vals)
    tag = "OK " if (ok and okv) else "FAIL"
    if not (ok and okv):
        fails += 1
        bad = np.nonzero(rk != exp_k)[0]
        print(f"   first key mismatch at {bad[:5]} of {bad.size}" if bad.size else "   values mismatch")
    print(f"{tag} {path} n={n} bits={bits} vb={vb} dist={dist}", flush=True)


sizes = [1, 5, 1000, 8192, 8193, 100003, (1 << 20) + 3]
if len(sys.argv) > 1:
    sizes = [int(x) for x in sys.argv[1:]]
for path in ("lsb", "msb"):
    for bits in (32, 64):
        for vb in (0, 4, 8):
            for n in sizes:
                for dist in ("uniform",) if vb == 8 else ("uniform", "lowent", "const", "few"):
                    try:
                        run_case(path, n, bits, vb, dist)
                    except Exception as e:  # noqa
                        fails += 1
                        print(f"EXC  {path} n={n} bits={bits} vb={vb} dist={dist}: {e!r}", flush=True)
                        if "CUDA" in repr(e) or "cuda" in repr(e):
                            print("fatal CUDA error, stopping"); print(json.dumps({"fails": fails})); sys.exit(1)
print(json.dumps({"fails": fails}))
sys.exit(1 if fails else 0)
