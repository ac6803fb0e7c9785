#!/usr/bin/env python
"""The reference LSB driver's own data (lsb/sort.cu:125-131): float keys uniform in (0,1], random u32 values.  Kernel breakdown."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_sort_b200 as gs
n = 1 << 28
g = torch.Generator(device="cuda"); g.manual_seed(0)
src = 1.0 - torch.rand(n, device="cuda", generator=g, dtype=torch.float32)
vsrc = torch.randint(-2**31, 2**31 - 1, (n,), device="cuda", dtype=torch.int32, generator=g)
for pairs, desc in ((True, False), (False, True), (False, False)):
    k0, k1 = torch.empty_like(src), torch.empty_like(src)
    v0 = torch.empty_like(vsrc) if pairs else None; v1 = torch.empty_like(vsrc) if pairs else None
    dk = gs.DoubleBuffer(k0, k1); dv = gs.DoubleBuffer(v0, v1) if pairs else None
    tb = gs.DeviceRadixSort._run(None, dk, dv, n, 0, None, desc, None, gs.KEY_F32)
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(5):
        k0.copy_(src)
        if pairs: v0.copy_(vsrc)
        dk = gs.DoubleBuffer(k0, k1); dv = gs.DoubleBuffer(v0, v1) if pairs else None
        if it == 4: gs.prof_enable(True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); gs.DeviceRadixSort._run(temp, dk, dv, n, 0, None, desc, None, gs.KEY_F32); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    rep = gs.prof_report(); gs.prof_enable(False)
    r = dk.Current()
    ok = bool((r[1:] >= r[:-1]).all().item()) if not desc else bool((r[1:] <= r[:-1]).all().item())
    print(json.dumps({"pairs": pairs, "descending": desc, "ms": [round(t, 3) for t in ts], "sorted": ok, "kernels": {k: (c, round(v, 3)) for k, (c, v) in rep.items()}}))

# the reference driver's own functions on the same data (oracle/_ref/libref_lsb.so = lsb/sort.cu compiled unmodified against toolkit CUB 2.8.2)
import ctypes
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_lsb.so")
if os.path.exists(so):
    lib = ctypes.CDLL(so)
    lib.ref_lsb_sortPairsGPU.restype = ctypes.c_float; lib.ref_lsb_sortKeysGPU.restype = ctypes.c_float
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    k0, k1 = torch.empty_like(src), torch.empty_like(src); v0, v1 = torch.empty_like(vsrc), torch.empty_like(vsrc)
    kv, kk = [], []
    for it in range(5):
        k0.copy_(src); v0.copy_(vsrc); torch.cuda.synchronize()
        kv.append(lib.ref_lsb_sortPairsGPU(P(k0), P(k1), P(v0), P(v1), ctypes.c_int(n)))
        k0.copy_(src); torch.cuda.synchronize()
        kk.append(lib.ref_lsb_sortKeysGPU(P(k0), P(k1), ctypes.c_int(n)))
    print(json.dumps({"reference_lsb_driver_cub_2.8.2": {"time_sort_kv_gpu_ms": [round(x, 3) for x in kv], "time_sort_k_gpu_ms (SortKeysDescending)": [round(x, 3) for x in kk]}}))
