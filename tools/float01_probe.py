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
