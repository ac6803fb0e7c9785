#!/usr/bin/env python
"""Multi-GPU correctness + timing of the exchange-as-level-0 sort (run under torchrun on N GPUs).
usage: torchrun --nproc-per-node N tools/dist_check.py [logn=24] [pairs=1] [dist=uniform] [reps=5]"""
import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_sort_b200 as gs
from gpu_sort_b200 import dist as gd

def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dname = sys.argv[3] if len(sys.argv) > 3 else "uniform"
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_l = 1 << logn; total = n_l * world
    src = torch.empty(n_l, dtype=torch.int32, device="cuda")
    gs.generate_keys(src, seed=0, dist=dname, param=3 if dname == "entropy" else 0, start=rank * n_l, total=total)
    vsrc = gs.iota(torch.empty(n_l, dtype=torch.int32, device="cuda"), start=rank * n_l) if pairs else None
    keys = torch.empty_like(src); vals = torch.empty_like(vsrc) if pairs else None
    din = gs.check(src, vsrc, key_type=gs.KEY_U32)[0]
    sorter = gd.ExchangeSorter(n_l, torch.int32, torch.int32 if pairs else None, key_type=gs.KEY_U32)
    times = []
    for it in range(reps + 2):
        keys.copy_(src)
        if pairs: vals.copy_(vsrc)
        torch.cuda.synchronize(); dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sorter.sort(keys, vals); e1.record(); torch.cuda.synchronize()
        if it >= 2: times.append(e0.elapsed_time(e1))
    k, v, info = sorter.result()
    keys.copy_(src)
    if pairs: vals.copy_(vsrc)
    sorter.sort(keys, vals, profile=True); prof = sorter.profile
    k, v, info = sorter.result()
    s, x, bad, vbad = gs.check(k, v, key_type=gs.KEY_U32) if k.numel() else (0, 0, 0, 0)
    mine = k.view(torch.int32)
    rec = {"sum": s, "bad": bad, "vbad": vbad if pairs else 0, "n": mine.numel(), "lo": int(mine[0].item()) & 0xFFFFFFFF if mine.numel() else None,
           "hi": int(mine[-1].item()) & 0xFFFFFFFF if mine.numel() else None, "din": din, "ms": float(np.median(times)), "path": info.get("path")}
    recs = [None] * world
    dist.all_gather_object(recs, rec)
    if rank == 0:
        ok = all(r["bad"] == 0 and r["vbad"] == 0 for r in recs) and sum(r["n"] for r in recs) == total
        nz = [r for r in recs if r["n"]]
        ok = ok and all(nz[i]["hi"] <= nz[i + 1]["lo"] for i in range(len(nz) - 1))
        ok = ok and sum(r["sum"] for r in recs) % (1 << 64) == sum(r["din"] for r in recs) % (1 << 64)
        ms = max(r["ms"] for r in recs)
        print(json.dumps({"world": world, "logn_per_gpu": logn, "pairs": pairs, "dist": dname, "ok": bool(ok), "ms_median_max_rank": round(ms, 3),
                          "gitems_s": round(total / ms * 1e-6, 2), "counts": [r["n"] for r in recs], "paths": sorted(set(r["path"] for r in recs)),
                          "imbalance": info.get("imbalance"), "phases_ms": prof["phases_ms"], "kernels_ms": prof["kernels_ms"]}), flush=True)
    dist.barrier(); dist.destroy_process_group()
main()
