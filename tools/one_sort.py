#!/usr/bin/env python
"""Two sorts of one bench workload (1 warm-up + 1), nothing else on the GPU: the ncu target (tools/gpu_round.sh).
usage: one_sort.py cfg2|cfg3|cfg4 [reps=2]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_sort_b200 as gs
import bench

def main():
    logn, kbits, vb, path, dist, param, S, desc = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    n = 1 << logn
    kt = gs.KEY_U32 if kbits == 32 else gs.KEY_U64
    kdt = torch.int32 if kbits == 32 else torch.int64
    src = torch.empty(n, dtype=kdt, device="cuda"); gs.generate_keys(src, seed=0, dist=dist, param=param)
    vsrc = gs.iota(torch.empty(n, dtype=torch.int32, device="cuda")) if vb else None
    k0, k1 = torch.empty_like(src), torch.empty_like(src)
    v0 = torch.empty_like(vsrc) if vb else None; v1 = torch.empty_like(vsrc) if vb else None
    if path == "lsb":
        tb = gs.DeviceRadixSort._run(None, gs.DoubleBuffer(k0, k1), gs.DoubleBuffer(v0, v1) if vb else None, n, 0, None, False, None, kt)
    else:
        tb = gs.rdxsrt_workspace_bytes(n, kt, vb)
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    for _ in range(reps):
        k0.copy_(src)
        if vb: v0.copy_(vsrc)
        if path == "lsb":
            gs.DeviceRadixSort._run(temp, gs.DoubleBuffer(k0, k1), gs.DoubleBuffer(v0, v1) if vb else None, n, 0, None, False, None, kt)
        else:
            gs.rdxsrt_unstable_sort(k0, v0, n, k1, v1, workspace=temp, key_type=kt)
    torch.cuda.synchronize()
    print("ok")
main()
