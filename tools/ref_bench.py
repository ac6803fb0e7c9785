#!/usr/bin/env python
"""Times the UNMODIFIED reference GPU builds (oracle/_ref/libref_{msb,lsb}.so) on the GPU box.

Reported baselines only: the reference MSB hybrid sort recompiled for sm_100a
(msb/src/sort/gpu_radix_sort.h:187-507) and the reference LSB driver's cub::DeviceRadixSort call
(lsb/sort.cu:25-76, toolkit CUB 2.8.2).  Prints one JSON line per case.  torch is used for device memory only.
"""
import ctypes, json, os, sys, time
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def gen_keys(n, bits, seed=0):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    if bits == 32:
        return torch.randint(-2**31, 2**31, (n,), dtype=torch.int32, device="cuda", generator=g)
    return torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)


def check_sorted_unsigned(t, bits):
    # compare as unsigned: flip the sign bit and compare signed
    flip = t ^ (-(2 ** (bits - 1)))
    return bool((flip[1:] >= flip[:-1]).all().item())


def time_call(fn, restore, reps):
    ts = []
    for _ in range(reps):
        restore()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    n = 1 << logn
    out = []
    msb = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libref_msb.so"))
    lsb = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libref_lsb.so"))
    msb.ref_msb_sort_device.restype = ctypes.c_int
    lsb.ref_lsb_cub_sort.restype = ctypes.c_int
    for bits, vb in ((32, 0), (32, 4), (64, 0)):
        if bits == 64 and logn > 27:
            n_eff = n  # 2^28 u64 = 2 GiB per buffer, fine
        else:
            n_eff = n
        src = gen_keys(n_eff, bits)
        k0 = torch.empty_like(src); k1 = torch.empty_like(src)
        vsrc = torch.arange(n_eff, dtype=torch.int32, device="cuda") if vb else None
        v0 = torch.empty_like(vsrc) if vb else None
        v1 = torch.empty_like(vsrc) if vb else None

        def restore():
            k0.copy_(src)
            if vb: v0.copy_(vsrc)

        # ---- reference MSB (allocations + stream creation are inside the call, as shipped)
        ok = ctypes.c_void_p(0); ov = ctypes.c_void_p(0)

        def run_msb():
            msb.ref_msb_sort_device(_ptr(k0), _ptr(v0), ctypes.c_ulonglong(n_eff), _ptr(k1), _ptr(v1),
                                    ctypes.c_int(bits), ctypes.c_int(vb), ctypes.byref(ok), ctypes.byref(ov))
        try:
            restore(); run_msb(); torch.cuda.synchronize()   # warm-up (first call prints thresholds)
            res = k0 if ok.value == k0.data_ptr() else k1
            good = check_sorted_unsigned(res, bits)
            med, best = time_call(run_msb, restore, reps)
            out.append({"impl": "reference-msb", "key_bits": bits, "value_bytes": vb, "n": n_eff, "ms_median": med,
                        "ms_best": best, "gkeys_s": n_eff / med * 1e-6, "sorted": good})
        except Exception as e:  # noqa
            out.append({"impl": "reference-msb", "key_bits": bits, "value_bytes": vb, "error": repr(e)})
        print(json.dumps(out[-1]), flush=True)

        # ---- reference LSB call shape: cub::DeviceRadixSort with DoubleBuffer, temp pre-allocated
        kt = 0 if bits == 32 else 1
        tb = ctypes.c_size_t(0); sel = ctypes.c_int(0)
        lsb.ref_lsb_cub_sort(None, ctypes.byref(tb), _ptr(k0), _ptr(k1), _ptr(v0), _ptr(v1), ctypes.c_int(n_eff),
                             ctypes.c_int(kt), ctypes.c_int(vb), 0, 0, bits, ctypes.byref(sel))
        temp = torch.empty(max(tb.value, 1), dtype=torch.uint8, device="cuda")

        def run_lsb():
            lsb.ref_lsb_cub_sort(_ptr(temp), ctypes.byref(tb), _ptr(k0), _ptr(k1), _ptr(v0), _ptr(v1),
                                 ctypes.c_int(n_eff), ctypes.c_int(kt), ctypes.c_int(vb), 0, 0, bits, ctypes.byref(sel))
        restore(); run_lsb(); torch.cuda.synchronize()
        res = k1 if sel.value else k0
        good = check_sorted_unsigned(res, bits)
        med, best = time_call(run_lsb, restore, reps)
        out.append({"impl": "reference-lsb-cub2.8.2", "key_bits": bits, "value_bytes": vb, "n": n_eff, "ms_median": med,
                    "ms_best": best, "gkeys_s": n_eff / med * 1e-6, "sorted": good, "temp_bytes": tb.value})
        print(json.dumps(out[-1]), flush=True)
        del src, k0, k1, vsrc, v0, v1, temp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
