#!/usr/bin/env python
"""Device-resident timing of the sort entry points (development tool; bench.py is the contract benchmark).

usage: perf.py [logn=28] [reps=7] [cases=msb32,lsb32,lsb32v4,msb32v4,msb64,lsb64] [dist=uniform] [param=0]
"""
import json, os, sys
import torch
sys.path.insert(0, ".")
import gpu_sort_b200 as gs


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    cases = (sys.argv[3] if len(sys.argv) > 3 else "msb32,lsb32,lsb32v4,msb32v4,msb64,lsb64").split(",")
    dist = sys.argv[4] if len(sys.argv) > 4 else "uniform"
    param = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    n = 1 << logn
    for case in cases:
        path = case[:3]; bits = int(case[3:5]); vb = int(case[6:]) if "v" in case[5:] else 0
        kt = gs.KEY_U32 if bits == 32 else gs.KEY_U64
        kdt = torch.int32 if bits == 32 else torch.int64
        vdt = torch.int32 if vb == 4 else torch.int64
        src = torch.empty(n, dtype=kdt, device="cuda"); gs.generate_keys(src, seed=0, dist=dist, param=param)
        vsrc = gs.iota(torch.empty(n, dtype=vdt, device="cuda")) if vb else None
        k0 = torch.empty_like(src); k1 = torch.empty_like(src)
        v0 = torch.empty_like(vsrc) if vb else None; v1 = torch.empty_like(vsrc) if vb else None
        ref_digest = gs.check(src, vsrc, key_type=kt)[:2]
        if path == "lsb":
            dk = gs.DoubleBuffer(k0, k1); dv = gs.DoubleBuffer(v0, v1) if vb else None
            tb = gs.DeviceRadixSort._run(None, dk, dv, n, 0, None, False, None, kt)
        else:
            tb = gs.rdxsrt_workspace_bytes(n, kt, vb)
        temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
        times = []
        res_k = res_v = None
        PROF_REPS = 4 if os.environ.get("PERF_PROF") else 0
        for it in range(reps + 2 + PROF_REPS):
            k0.copy_(src)
            if vb: v0.copy_(vsrc)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            if path == "lsb":
                dk = gs.DoubleBuffer(k0, k1); dv = gs.DoubleBuffer(v0, v1) if vb else None
                gs.DeviceRadixSort._run(temp, dk, dv, n, 0, None, False, None, kt)
                res_k = dk.Current(); res_v = dv.Current() if vb else None
            else:
                r = gs.rdxsrt_unstable_sort(k0, v0, n, k1, v1, workspace=temp, key_type=kt)
                res_k = r.sorted_keys; res_v = r.sorted_values
            e1.record(); torch.cuda.synchronize()
            if it >= 2 and it < reps + 2: times.append(e0.elapsed_time(e1))
            if it == reps + 1 and os.environ.get("PERF_PROF"): gs.prof_enable(True)      # PROF_REPS more repetitions, timed per kernel family
        prof = None
        if os.environ.get("PERF_PROF"):
            prof = {k: [v[0] // PROF_REPS, round(v[1] / PROF_REPS, 4)] for k, v in gs.prof_report().items()}; gs.prof_enable(False)
        times.sort()
        s, x, bad, vbad = gs.check(res_k, res_v, key_type=kt)
        med = times[len(times) // 2]
        print(json.dumps({"case": case, "dist": dist, "param": param, "n": n, "ms_median": round(med, 4), "ms_best": round(times[0], 4),
                          "gkeys_s": round(n / med * 1e-6, 2), "temp_mb": round(tb / 2**20, 1), "sorted": bad == 0,
                          "multiset_ok": (s, x) == ref_digest, "stable_iota": vbad == 0 if vb else None, "prof": prof}), flush=True)
        del src, vsrc, k0, k1, v0, v1, temp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
