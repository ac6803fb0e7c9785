#!/usr/bin/env python
"""Store bandwidth into a PEER GPU's memory over NVLink vs. local HBM, by access pattern (run under torchrun, 2 GPUs)."""
import os, sys, json, ctypes
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_sort_b200 as gs
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 1 << 30
buf = symm.empty(nbytes // 4, dtype=torch.int32, device=f"cuda:{local}")
h = symm.rendezvous(buf, dist.group.WORLD.group_name)
ptrs = [int(p) for p in h.buffer_ptrs]
f = gs.lib.b200_util_store_probe
f.restype = ctypes.c_int; f.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
res = []
for target in ("local", "peer"):
    dst = ptrs[rank] if target == "local" else ptrs[(rank + 1) % world]
    for mode, chunk in ((0, 0), (1, 0), (2, 64), (2, 128), (2, 256), (2, 512), (2, 2048), (3, 64), (3, 128), (3, 256), (3, 512), (3, 2048), (4, 128), (4, 1024), (4, 4096), (2, 1024), (2, 4096)):
        for grid, block in ((148 * 2, 512),):
            dist.barrier(); torch.cuda.synchronize()
            ts = []
            for it in range(4):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); f(dst, nbytes, mode, max(chunk, 16), grid, block, None); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            if rank == 0:
                res.append({"target": target, "mode": mode, "chunk": chunk, "grid": grid, "block": block, "GBps": round(nbytes / min(ts[1:]) * 1e-6, 1)})
if rank == 0:
    for r in res: print(json.dumps(r))
dist.barrier(); dist.destroy_process_group()
