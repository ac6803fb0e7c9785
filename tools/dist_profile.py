#!/usr/bin/env python
"""Per-phase CUDA-event timing of the multi-GPU sort (run under torchrun): histogram, all-reduce, splitters, range partition,
count all-gather, all-to-all, local sort.  usage: torchrun ... tools/dist_profile.py [logn=28] [pairs=0]"""
import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_sort_b200 as gs
from gpu_sort_b200 import dist as gd

def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << logn
    src = torch.empty(n, dtype=torch.int32, device="cuda"); gs.generate_keys(src, seed=0, start=rank * n, total=n * world)
    vsrc = gs.iota(torch.empty(n, dtype=torch.int32, device="cuda"), start=rank * n) if pairs else None
    ops = gd.CudaOps(gs.KEY_U32)
    def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
    for it in range(4):
        keys = src.clone(); vals = vsrc.clone() if pairs else None
        torch.cuda.synchronize(); dist.barrier()
        t = {}
        e0 = ev(); counts = ops.histogram(keys, 14); e1 = ev()
        g = counts.clone(); dist.all_reduce(g); e2 = ev()
        w0 = time.time(); sp = gd.choose_splitters(g, world); w1 = time.time(); e3 = ev()
        pk, pv, offs = ops.partition(keys, vals, 14, sp, counts); e4 = ev()
        sc = (offs[1:] - offs[:-1]).contiguous(); gl = [torch.empty_like(sc) for _ in range(world)]; dist.all_gather(gl, sc)
        m = np.stack([x.cpu().numpy() for x in gl]); send, recv, nr = gd.receive_layout(m, rank); e5 = ev()
        rk = torch.empty(nr, dtype=torch.int32, device="cuda"); rv = torch.empty(nr, dtype=torch.int32, device="cuda") if pairs else None
        e5b = ev()
        dist.all_to_all_single(rk, pk, output_split_sizes=recv, input_split_sizes=send)
        if pairs: dist.all_to_all_single(rv, pv, output_split_sizes=recv, input_split_sizes=send)
        e6 = ev()
        sk, sv = ops.local_sort(rk, rv, nr, bool(pairs)); e7 = ev()
        torch.cuda.synchronize()
        if rank == 0 and it >= 1:
            print(json.dumps({"hist": e0.elapsed_time(e1), "allreduce": e1.elapsed_time(e2), "splitters_gpu_gap": e2.elapsed_time(e3), "splitters_host_s": w1 - w0,
                              "partition": e3.elapsed_time(e4), "allgather+d2h": e4.elapsed_time(e5), "alloc": e5.elapsed_time(e5b), "all_to_all": e5b.elapsed_time(e6),
                              "local_sort": e6.elapsed_time(e7), "total": e0.elapsed_time(e7), "send": send}), flush=True)
    dist.barrier(); dist.destroy_process_group()
main()
