#!/usr/bin/env python
"""bench.py -- the contract benchmark of the B200-native radix sort (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1|cfg2|cfg3|cfg4|cfg5] [--logn L]

Metric (BASELINE.json): Gkeys/s.  A step = ONE sort of one batch of synthetic keys.
  N = 1 (default) : BASELINE config 2 -- 2^28 uniform uint32 keys, keys-only, MSB hybrid sort (rdxsrt_unstable_sort call
                    shape) through the C ABI.  `value` times the sort call alone with the keys resident in HBM (CUDA events
                    around every step, the input is restored by a D2D copy between steps -- the reference's own protocol,
                    lsb/sort.cu:141-146); `e2e` times the host-pointer entry point (b200_msb_sort_host: H2D + sort + D2H).
  N > 1 (torchrun): weak scaling -- every rank holds 2^28 keys of ONE global array of N*2^28 keys, sorted with the multi-GPU
                    path (histogram all-reduce, splitters, key all-to-all over NVLink, local sort): gpu_sort_b200/dist.py.
  --impl reference: the UNMODIFIED reference (oracle/_ref/libref_msb.so, compiled from /root/reference) on the same
                    workload on the same GPU; if that library is missing, the CPU port (oracle/) on a bounded sample.
One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (log2 n, key bits, value bytes, path, dist, param, algorithmic full sweeps S, description)
    "cfg1": (24, 32, 0, "msb", "uniform", 0, 3, "2^24 uniform uint32 keys, keys-only, MSB hybrid"),
    "cfg2": (28, 32, 0, "msb", "uniform", 0, 3, "2^28 uniform uint32 keys, keys-only, MSB hybrid radix sort"),
    "cfg3": (28, 32, 4, "lsb", "uniform", 0, 3, "2^28 uint32 key + uint32 value pairs, stable sort (cub::DeviceRadixSort call shape)"),
    "cfg4": (29, 64, 0, "msb", "zipf_hash", 0, 3, "2^29 uint64 keys, Zipf-skewed, MSB hybrid"),
    # one GPU's share of BASELINE config 5 (2^32 pairs over 8 GPUs): the single-GPU / reference-arm form of the multi-GPU workload
    "cfg5": (29, 32, 4, "lsb", "uniform", 0, 4, "2^29 uint32 key + uint32 value pairs (one GPU's share of config 5), stable sort"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_baseline(n_sample=1 << 26):
    """std::sort on one core and __gnu_parallel::sort on all cores (oracle/libcpusort.so) over a bounded sample of the
    workload's keys -- a reported baseline, not the target."""
    so = os.path.join(ROOT, "oracle", "libcpusort.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libcpusort.so"], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    lib.cpu_sort_u32.restype = ctypes.c_double
    lib.cpu_sort_u32.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
    lib.cpu_sort_max_threads.restype = ctypes.c_int
    from tests import oracle_lib
    keys = oracle_lib.load().gen_keys(n_sample, 32, seed=0, dist="uniform")
    cores = lib.cpu_sort_max_threads()
    a = keys.copy()
    t_par = lib.cpu_sort_u32(a.ctypes.data, a.size, 0)
    n1 = n_sample >> 2
    b = keys[:n1].copy()
    t_one = lib.cpu_sort_u32(b.ctypes.data, b.size, 1)
    assert np.all(a[1:] >= a[:-1])
    return {"value": round(n_sample / t_par / 1e9, 4), "unit": "Gkeys/s", "cores": cores, "kind": "port",
            "sample": f"__gnu_parallel::sort of 2^{int(np.log2(n_sample))} uniform u32 keys (same generator/seed as the workload) on {cores} threads",
            "std_sort_1core_gkeys_s": round(n1 / t_one / 1e9, 4), "std_sort_sample": f"2^{int(np.log2(n1))} keys"}


def time_steps(step_fn, restore_fn, steps, warmup, barrier):
    """W untimed warm-ups, then K steps; each step = restore (untimed) + sort between two CUDA events on the current stream."""
    for _ in range(warmup):
        restore_fn(); step_fn()
    torch.cuda.synchronize(); barrier()
    evs = []
    wall0 = time.time()
    for _ in range(steps):
        restore_fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step_fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize(); barrier()
    wall = time.time() - wall0
    ms = [a.elapsed_time(b) for a, b in evs]
    return ms, wall


def run_ours_single(args, wl):
    import gpu_sort_b200 as gs
    logn, kbits, vb, path, dist, param, S, desc = wl
    if args.logn:
        logn = args.logn
    n = 1 << logn
    kt = gs.KEY_U32 if kbits == 32 else gs.KEY_U64
    kdt = torch.int32 if kbits == 32 else torch.int64
    vdt = torch.int32 if vb == 4 else torch.int64
    src = torch.empty(n, dtype=kdt, device="cuda"); gs.generate_keys(src, seed=0, dist=dist, param=param)
    vsrc = gs.iota(torch.empty(n, dtype=vdt, device="cuda")) if vb else None
    k0, k1 = torch.empty_like(src), torch.empty_like(src)
    v0 = torch.empty_like(vsrc) if vb else None; v1 = torch.empty_like(vsrc) if vb else None
    digest_in = gs.check(src, vsrc, key_type=kt)[:2]
    if path == "lsb":
        tb = gs.DeviceRadixSort._run(None, gs.DoubleBuffer(k0, k1), gs.DoubleBuffer(v0, v1) if vb else None, n, 0, None, False, None, kt)
    else:
        tb = gs.rdxsrt_workspace_bytes(n, kt, vb)
    temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
    res = {}

    def restore():
        k0.copy_(src)
        if vb:
            v0.copy_(vsrc)

    def step():
        if path == "lsb":
            dk = gs.DoubleBuffer(k0, k1); dv = gs.DoubleBuffer(v0, v1) if vb else None
            gs.DeviceRadixSort._run(temp, dk, dv, n, 0, None, False, None, kt)
            res["k"], res["v"] = dk.Current(), (dv.Current() if vb else None)
        else:
            r = gs.rdxsrt_unstable_sort(k0, v0 if vb else None, n, k1, v1 if vb else None, workspace=temp, key_type=kt)
            res["k"], res["v"] = r.sorted_keys, r.sorted_values

    clocks = ClockSampler(torch.cuda.current_device()); clocks.start()
    ms, wall = time_steps(step, restore, args.steps, args.warmup, lambda: None)
    clk = clocks.stop()
    s, x, bad, vbad = gs.check(res["k"], res["v"], key_type=kt)
    ok = bad == 0 and (s, x) == digest_in and (path != "lsb" or not vb or vbad == 0)

    # ---- per-kernel durations: the same steps again with every launch bracketed by CUDA events (b200_prof_*)
    gs.prof_enable(True)
    for _ in range(args.steps):
        restore(); step()
    torch.cuda.synchronize()
    launches_total = gs.prof_launches()          # every launch site of the library counts itself while profiling is on
    prof = gs.prof_report(); gs.prof_enable(False)
    launches_per_step = launches_total / args.steps
    dom = max(prof.items(), key=lambda kv: kv[1][1])
    dom_name, (dom_cnt, dom_ms) = dom
    kbytes, vbytes = kbits // 8, vb
    sweep_bytes = 2 * n * (kbytes + vbytes)
    # algorithmic bytes of the dominant kernel family per step (DESIGN.md "Kernels"): S - 1 scatter sweeps each read and
    # write every key (+ value) once; each is preceded by one histogram read of the keys; the on-chip sort reads and writes
    # every key (+ value) once.  Both entry points run the same MSD engine (stable or not).
    fam_bytes = {"scatter": (S - 1) * sweep_bytes, "scatter_stable": (S - 1) * sweep_bytes, "scatter_onesweep": S * sweep_bytes,
                 "tile_hist": (S - 1) * n * kbytes, "hist_all": n * kbytes, "local_sort_lsd": sweep_bytes, "local_sort_count": sweep_bytes,
                 "local_sort_rank": sweep_bytes, "local_sort_bitmap": sweep_bytes}
    dom_bytes = fam_bytes.get(dom_name, sweep_bytes)
    dom_ms_per_step = dom_ms / args.steps
    peak, peak_src = peaks()
    achieved = dom_bytes / (dom_ms_per_step * 1e-3) / 1e9
    med = float(np.median(ms)); mean = float(np.mean(ms))
    whole_bytes = n * kbytes + S * sweep_bytes       # the contract figure of SURVEY.md section 8(d): one histogram read + S sweeps
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_step": dom_bytes, "kernel_ms_per_step": round(dom_ms_per_step, 4),
                "kernel_launches_per_step": dom_cnt / args.steps, "kernel_share_of_step": round(dom_ms_per_step / (sum(v for _, v in prof.values()) / args.steps), 3),
                "whole_sort": {"algorithmic_bytes": whole_bytes, "bytes_per_key": whole_bytes / n, "achieved_gbs": round(whole_bytes / (mean * 1e-3) / 1e9, 1),
                               "frac": round(whole_bytes / (mean * 1e-3) / 1e9 / peak, 4)},
                "kernels_ms_per_step": {k: round(v / args.steps, 4) for k, (c, v) in prof.items()}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(dom_name, {}).get(args.workload)
        except Exception:
            pass

    # ---- end to end through the host-pointer entry point (pinned host buffers; H2D + sort + D2H inside the timed region)
    e2e = None
    if path == "msb" and not args.no_e2e:
        hk = torch.empty(n, dtype=kdt).pin_memory(); hk.copy_(src)
        ho = torch.empty(n, dtype=kdt).pin_memory()
        hv = hvo = None
        if vb:
            hv = torch.empty(n, dtype=vdt).pin_memory(); hv.copy_(vsrc); hvo = torch.empty(n, dtype=vdt).pin_memory()
        e2e_steps = max(3, min(args.steps, 10))
        call = lambda: gs._check(gs.lib.b200_msb_sort_host(hk.data_ptr(), hv.data_ptr() if vb else None, n, ho.data_ptr(), hvo.data_ptr() if vb else None, kt, vb), "b200_msb_sort_host")
        call(); call()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()                       # synchronous on return, like the reference's wrapper (gpu_radix_sort.h:510-541)
        dt = (time.perf_counter() - t0) / e2e_steps
        a = ho.numpy().view(np.uint32 if kbits == 32 else np.uint64)
        ok = ok and bool(np.all(a[1:] >= a[:-1]))
        e2e = {"value": round(n / dt / 1e9, 3), "unit": "Gkeys/s", "h2d_bytes_per_step": n * (kbytes + vbytes), "d2h_bytes_per_step": n * (kbytes + vbytes),
               "ms_per_step": round(dt * 1e3, 3), "steps": e2e_steps, "entry": "b200_msb_sort_host (pinned host buffers)"}

    line = {"metric": "Gkeys/s", "value": round(n / (mean * 1e-3) / 1e9, 3), "unit": "Gkeys/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(mean, 4), "ms_median": round(med, 4), "value_median": round(n / (med * 1e-3) / 1e9, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32" if kbits == 32 else "u64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "n": n, "path": path, "key_bits": kbits, "value_bytes": vb, "dist": dist,
                       "l2": "inputs (>= 1 GiB) larger than the 126 MB L2; input restored by an untimed D2D copy between steps",
                       "timing": "CUDA events around each sort call on the launching stream, summed over the K steps"},
            "roofline": roofline, "clocks": clk, "gpu_launches": int(round(launches_per_step * args.steps)), "gpu_launches_per_step": launches_per_step,
            "verified": ok, "wall_s_timed_region": round(wall, 3)}
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline()
    return line


def extra_workloads(args):
    """BASELINE's metric reads "keys/pairs": the default line also carries the pairs config (cfg3) and the skewed 64-bit config
    (cfg4, Zipf-hashed keys) as short runs with their own roofline figures (full lines: bench.py --workload cfg3 | cfg4), and the
    single-GPU point of the multi-GPU weak-scaling series (cfg5)."""
    out = {}
    # cfg5 here = ONE GPU's share of the multi-GPU config (2^29 pairs, stable): the N = 1 point of the weak-scaling series that
    # `bench.py --gpus N` measures on N > 1 GPUs (those lines carry the same figure as `single_gpu_same_workload`)
    for name in ("cfg3", "cfg4", "cfg5"):
        sub = argparse.Namespace(**vars(args))
        sub.workload = name; sub.steps = max(3, min(args.steps, 5)); sub.warmup = 3; sub.no_cpu = True; sub.no_e2e = True; sub.logn = 0
        try:
            l = run_ours_single(sub, WORKLOADS[name])
            out[name] = {k: l[k] for k in ("value", "unit", "ms_per_step", "ms_median", "value_median", "steps", "dtype", "config", "roofline", "gpu_launches_per_step", "verified")}
        except Exception as e:          # the headline line must not die with an extra
            out[name] = {"error": repr(e)}
        torch.cuda.empty_cache()
    return out


def run_ours_multi(args, rank, world):
    """BASELINE config 5, weak-scaled: every rank holds 2^logn (default 2^29) uint32 key + uint32 value pairs of ONE global array
    (N = 8 -> 2^32 pairs), sorted with the multi-GPU path of gpu_sort_b200/dist.py (ExchangeSorter: per-tile digit counts,
    256-bin count all-gather, stable scatter straight into the peers' receive buffers over NVLink, segmented finish)."""
    import torch.distributed as dist
    import gpu_sort_b200 as gs
    from gpu_sort_b200 import dist as gd
    pairs = args.workload != "cfg2"
    logn = args.logn or (29 if pairs else 28)
    n_l = 1 << logn
    total = n_l * world
    src = torch.empty(n_l, dtype=torch.int32, device="cuda")
    gs.generate_keys(src, seed=0, dist="uniform", start=rank * n_l, total=total)
    vsrc = gs.iota(torch.empty(n_l, dtype=torch.int32, device="cuda"), start=rank * n_l) if pairs else None
    keys = torch.empty_like(src); vals = torch.empty_like(vsrc) if pairs else None
    din = gs.check(src, vsrc, key_type=gs.KEY_U32)[0]

    def restore():
        keys.copy_(src)
        if pairs:
            vals.copy_(vsrc)

    if args.nccl_exchange:
        sorter = gd.DistSorter(n_l, torch.int32, torch.int32 if pairs else None, key_type=gs.KEY_U32, fused=False)
        res = {}

        def step():
            res["out"] = sorter.sort(keys, vals)
        fetch = lambda: res["out"]
    else:
        sorter = gd.ExchangeSorter(n_l, torch.int32, torch.int32 if pairs else None, key_type=gs.KEY_U32)

        def step():
            sorter.sort(keys, vals)          # enqueue only: no host read-back inside the timed region
        fetch = sorter.result

    clocks = ClockSampler(torch.cuda.current_device()); clocks.start()
    ms, wall = time_steps(step, restore, args.steps, args.warmup, dist.barrier)
    clk = clocks.stop()
    rk, rv, info = fetch()
    t = torch.tensor([float(np.sum(ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                       # max over ranks of the device-timed K steps
    total_ms = float(t.item())
    tm = torch.tensor([float(np.median(ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    # validation: per-rank sortedness (+ stability), boundary order between neighbouring ranks, global multiset digest
    s, x, bad, vbad = gs.check(rk, rv, key_type=gs.KEY_U32) if rk.numel() else (0, 0, 0, 0)
    mine = rk.view(torch.int32)
    lo = int(mine[0].item()) & 0xFFFFFFFF if mine.numel() else None
    hi = int(mine[-1].item()) & 0xFFFFFFFF if mine.numel() else None
    rec = {"sum": s, "bad": bad, "vbad": vbad if pairs else 0, "n": mine.numel(), "lo": lo, "hi": hi, "din": din, "path": info.get("path", "key-range")}
    recs = [None] * world
    dist.all_gather_object(recs, rec)
    ok = all(r["bad"] == 0 and r["vbad"] == 0 for r in recs) and sum(r["n"] for r in recs) == total
    nz = [r for r in recs if r["n"]]
    ok = ok and all(nz[i]["hi"] <= nz[i + 1]["lo"] for i in range(len(nz) - 1))
    ok = ok and sum(r["sum"] for r in recs) % (1 << 64) == sum(r["din"] for r in recs) % (1 << 64)
    # ---- kernels launched per step and phase times: one more sort with every launch bracketed by CUDA events
    restore()
    phases, launches = {}, None
    if not args.nccl_exchange:
        sorter.sort(keys, vals, profile=True)
        phases = dict(sorter.profile["phases_ms"], kernels=sorter.profile["kernels_ms"])
        launches = sorter.profile.get("launches")
    # ---- the same per-GPU workload on ONE GPU (rank 0's shard, the single-GPU stable sort): the weak-scaling reference point
    local = None
    if rank == 0 and not args.no_local_ref:
        k0, k1 = torch.empty_like(src), torch.empty_like(src)
        v0 = torch.empty_like(vsrc) if pairs else None; v1 = torch.empty_like(vsrc) if pairs else None
        dk0 = gs.DoubleBuffer(k0, k1); dv0 = gs.DoubleBuffer(v0, v1) if pairs else None
        tb = gs.DeviceRadixSort._run(None, dk0, dv0, n_l, 0, None, False, None, gs.KEY_U32)
        temp = torch.empty(tb, dtype=torch.uint8, device="cuda")
        lt = []
        for it in range(5):
            k0.copy_(src)
            if pairs:
                v0.copy_(vsrc)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gs.DeviceRadixSort._run(temp, gs.DoubleBuffer(k0, k1), gs.DoubleBuffer(v0, v1) if pairs else None, n_l, 0, None, False, None, gs.KEY_U32)
            e1.record(); torch.cuda.synchronize()
            if it >= 2:
                lt.append(e0.elapsed_time(e1))
        lm = float(np.median(lt))
        local = {"ms": round(lm, 4), "value": round(n_l / (lm * 1e-3) / 1e9, 3), "unit": "Gkeys/s",
                 "what": f"single-GPU stable sort (b200_lsb_sort) of one rank's 2^{logn} items, same process, other ranks idle"}
        del k0, k1, v0, v1, temp
    dist.barrier()
    # ---- end to end: every rank's shard starts in pinned host memory and the sorted range it ends up owning returns there
    e2e = None
    if not args.no_e2e:
        hk = torch.empty(n_l, dtype=torch.int32).pin_memory(); hk.copy_(src)
        hv = None
        if pairs:
            hv = torch.empty(n_l, dtype=torch.int32).pin_memory(); hv.copy_(vsrc)
        ho = torch.empty(sorter.cap, dtype=torch.int32).pin_memory()
        hvo = torch.empty(sorter.cap, dtype=torch.int32).pin_memory() if pairs else None

        def e2e_step():
            keys.copy_(hk, non_blocking=True)
            if pairs:
                vals.copy_(hv, non_blocking=True)
            step()
            k, v, inf = fetch()
            ho[:inf["count"]].copy_(k, non_blocking=True)
            if pairs:
                hvo[:inf["count"]].copy_(v, non_blocking=True)
            torch.cuda.synchronize()
        e2e_step()
        dist.barrier()
        k_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        dist.barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / k_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        bpk = 8 if pairs else 4
        e2e = {"value": round(total / float(dt.item()) / 1e9, 3), "unit": "Gkeys/s", "h2d_bytes_per_step": total * bpk, "d2h_bytes_per_step": total * bpk,
               "ms_per_step": round(float(dt.item()) * 1e3, 3), "entry": "ExchangeSorter.sort + result with pinned host shards (H2D + sort + D2H per rank)"}
    if rank != 0:
        return None
    mean = total_ms / args.steps
    med = float(tm.item())
    kb = 8 if pairs else 4
    value = total / (mean * 1e-3) / 1e9
    nv_bytes = int(n_l * kb * (world - 1) / world)
    xms = phases.get("scatter")
    line = {"metric": "Gkeys/s", "value": round(value, 3), "unit": "Gkeys/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(mean, 4), "ms_median": round(med, 4), "value_median": round(total / (med * 1e-3) / 1e9, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": (f"cfg5: uint32 key + uint32 value pairs, 2^{logn} per GPU of one global array of {world}x2^{logn}" + (" (= 2^32 pairs, BASELINE config 5)" if total == 1 << 32 else "")
                                    if pairs else f"cfg2 weak-scaled: 2^{logn} uniform uint32 keys per GPU of one global array of {world}x2^{logn}")
                                   + "; per-tile digit counts + 256-bin count all-gather + stable scatter into the peers' receive buffers over NVLink + segmented finish",
                       "n_total": total, "n_per_gpu": n_l, "value_bytes": 4 if pairs else 0, "path": "nccl all_to_all baseline" if args.nccl_exchange else sorted(set(r["path"] for r in recs)),
                       "l2": "inputs larger than L2; restored by an untimed D2D copy between steps", "timing": "CUDA events per step on each rank; max over ranks of the K-step sum"},
            "exchange": {"imbalance": round(float(info.get("imbalance", 1.0)), 4), "nvlink_bytes_out_per_gpu": nv_bytes,
                         "nvlink_gbs_per_gpu_during_scatter": round(nv_bytes / (xms * 1e-3) / 1e9, 1) if xms else None, "phases_ms_rank0": phases},
            "single_gpu_same_workload": local,
            "weak_scaling_efficiency_vs_single_gpu_same_workload": round(value / (world * local["value"]), 3) if local else None,
            "clocks": clk, "gpu_launches": int(launches * args.steps) if launches else None, "gpu_launches_per_step": launches,
            "e2e": e2e, "verified": bool(ok), "wall_s_timed_region": round(wall, 3)}
    return line


def run_reference(args, wl):
    """The unmodified reference's GPU build (oracle/_ref) on the same workload; CPU port if it is not there."""
    logn, kbits, vb, path, dist, param, S, desc = wl
    if args.logn:
        logn = args.logn
    n = 1 << logn
    so = os.path.join(ROOT, "oracle", "_ref", "libref_msb.so" if path == "msb" else "libref_lsb.so")
    if not (os.path.exists(so) and torch.cuda.is_available()):
        cb = cpu_baseline()
        return {"impl": "reference", "metric": "Gkeys/s", "value": cb["value"], "unit": "Gkeys/s", "n_gpus": 1, "steps": 1, "warmup": 0, "higher_is_better": True,
                "config": {"workload": f"{args.workload}: {desc} (bounded CPU sample)"}, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "Gkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    # inputs and result check come from the CPU oracle and torch: this arm must not load libb200sort.so
    from tests import oracle_lib
    orc = oracle_lib.load()
    kdt = torch.int32 if kbits == 32 else torch.int64
    vdt = torch.int32 if vb == 4 else torch.int64
    keys_np = orc.gen_keys(n, kbits, seed=0, dist=dist, param=param)          # same generator and seed as the product arm (b200_util_generate_keys)
    src = torch.from_numpy(keys_np.view(np.int32 if kbits == 32 else np.int64)).cuda()
    vsrc = torch.arange(n, dtype=vdt, device="cuda") if vb else None
    k0, k1 = torch.empty_like(src), torch.empty_like(src)
    v0 = torch.empty_like(vsrc) if vb else None; v1 = torch.empty_like(vsrc) if vb else None
    P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
    lib = ctypes.CDLL(so)
    res = {}

    def restore():
        k0.copy_(src)
        if vb:
            v0.copy_(vsrc)

    if path == "msb":
        lib.ref_msb_sort_device.restype = ctypes.c_int
        ok_, ov_ = ctypes.c_void_p(0), ctypes.c_void_p(0)
        # the reference's own pre_allocated_dm parameter (gpu_radix_sort.h:196,224-228): its temporary memory is allocated once,
        # outside the timed call; the as-shipped call (eight cudaMalloc/cudaFree inside) is timed beside it
        prealloc = hasattr(lib, "ref_msb_sort_device_prealloc") and (kbits, vb) in ((32, 0), (64, 0), (32, 4), (64, 8))
        entry = lib.ref_msb_sort_device_prealloc if prealloc else lib.ref_msb_sort_device
        entry.restype = ctypes.c_int

        def call(fn):
            fn(P(k0), P(v0), ctypes.c_ulonglong(n), P(k1), P(v1), ctypes.c_int(kbits), ctypes.c_int(vb), ctypes.byref(ok_), ctypes.byref(ov_))
            res["k"] = k0 if ok_.value == k0.data_ptr() else k1
            res["v"] = (v0 if ov_.value == v0.data_ptr() else v1) if vb else None

        def step():
            call(entry)

        def step_as_shipped():
            call(lib.ref_msb_sort_device)
        api = ("rdxsrt_unstable_sort (msb/src/sort/gpu_radix_sort.h:187-507), unmodified, sm_100a build, "
               + ("pre_allocated_dm reused across calls" if prealloc else "as shipped (allocates inside the call)"))
    else:
        lib.ref_lsb_cub_sort.restype = ctypes.c_int
        tb = ctypes.c_size_t(0); sel = ctypes.c_int(0)
        rkt = 0 if kbits == 32 else 1
        lib.ref_lsb_cub_sort(None, ctypes.byref(tb), P(k0), P(k1), P(v0), P(v1), ctypes.c_int(n), rkt, vb, 0, 0, kbits, ctypes.byref(sel))
        temp = torch.empty(max(tb.value, 1), dtype=torch.uint8, device="cuda")

        def step():
            lib.ref_lsb_cub_sort(P(temp), ctypes.byref(tb), P(k0), P(k1), P(v0), P(v1), ctypes.c_int(n), rkt, vb, 0, 0, kbits, ctypes.byref(sel))
            res["k"] = k1 if sel.value else k0
            res["v"] = (v1 if sel.value else v0) if vb else None
        api = "cub::DeviceRadixSort call shape of lsb/sort.cu:25-76 (toolkit CUB 2.8.2; vendored 1.6.4 cannot assemble for sm_100)"

    clocks = ClockSampler(torch.cuda.current_device()); clocks.start()
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)       # the reference prints its thresholds on first use
    shipped = None
    try:
        ms, wall = time_steps(step, restore, args.steps, args.warmup, lambda: None)
        clk = clocks.stop()
        if path == "msb":       # the as-shipped call, timed without the nvidia-smi poller (its cudaMalloc/cudaFree contend with NVML queries)
            ms2, _ = time_steps(step_as_shipped, restore, max(3, min(args.steps, 10)), 1, lambda: None)
            shipped = {"ms_mean": round(float(np.mean(ms2)), 4), "ms_median": round(float(np.median(ms2)), 4),
                       "gkeys_s_median": round(n / (float(np.median(ms2)) * 1e-3) / 1e9, 3)}
            step()              # leave the result of the measured entry in place for the check below
    finally:
        os.dup2(saved, 1); os.close(devnull)
    out_k = res["k"].cpu().numpy().view(np.uint32 if kbits == 32 else np.uint64)
    bad = int(orc.count_unsorted(out_k, "u32" if kbits == 32 else "u64"))
    out_v = res["v"].cpu().numpy().view(np.uint32 if vb == 4 else np.uint64) if vb else None
    in_v = np.arange(n, dtype=np.uint32 if vb == 4 else np.uint64) if vb else None
    same_multiset = orc.digest(out_k, out_v) == orc.digest(keys_np, in_v)
    bad = bad if same_multiset else max(bad, 1)
    mean = float(np.mean(ms)); med = float(np.median(ms))
    line = {"impl": "reference", "metric": "Gkeys/s", "value": round(n / (mean * 1e-3) / 1e9, 3), "unit": "Gkeys/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(mean, 4), "ms_median": round(med, 4), "value_median": round(n / (med * 1e-3) / 1e9, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32" if kbits == 32 else "u64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "n": n, "path": path, "api": api}, "clocks": clk, "verified": bad == 0}
    if shipped is not None:
        line["as_shipped_no_prealloc"] = shipped
    # e2e: the reference's host-pointer wrapper (malloc + H2D + sort + D2H + free, gpu_radix_sort.h:510-541)
    if path == "msb" and vb == 0 and not args.no_e2e:
        lib.ref_msb_sort_keys_host.restype = ctypes.c_int
        hk = torch.empty(n, dtype=kdt).pin_memory(); hk.copy_(src)
        ho = torch.empty(n, dtype=kdt).pin_memory()
        call = lambda: lib.ref_msb_sort_keys_host(ctypes.c_void_p(hk.data_ptr()), ctypes.c_ulonglong(n), ctypes.c_void_p(ho.data_ptr()), ctypes.c_int(kbits))
        saved = os.dup(1); devnull = os.open(os.devnull, os.O_WRONLY); os.dup2(devnull, 1)
        try:
            call()
            k = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(k):
                call()
            dt = (time.perf_counter() - t0) / k
        finally:
            os.dup2(saved, 1); os.close(devnull)
        line["e2e"] = {"value": round(n / dt / 1e9, 3), "unit": "Gkeys/s", "h2d_bytes_per_step": n * kbits // 8, "d2h_bytes_per_step": n * kbits // 8,
                       "ms_per_step": round(dt * 1e3, 3), "entry": "rdxsrt_unstable_sort_keys (host pointers)"}
    else:
        line["e2e"] = {"value": line["value"], "unit": "Gkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS), help="default: cfg2 on one GPU, cfg5 on several")
    ap.add_argument("--logn", type=int, default=0, help="override log2(keys per GPU)")
    ap.add_argument("--dist", default=None, help="override the workload's key distribution: uniform|entropy|zipf_rank|zipf_hash|sorted|reverse|constant")
    ap.add_argument("--param", type=int, default=0, help="distribution parameter (entropy: number of AND rounds)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="single GPU: skip the short cfg3 / cfg4 runs carried in the default line")
    ap.add_argument("--no-local-ref", action="store_true", help="multi-GPU: skip the single-GPU reference point of the same per-GPU workload")
    ap.add_argument("--nccl-exchange", action="store_true", help="multi-GPU: NCCL all_to_all_single instead of the fused peer-memory scatter")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = "cfg2" if world == 1 else "cfg5"
    if args.dist is not None:       # the skew sweep of config 4: same sizes and types, another key distribution
        w = list(WORKLOADS[args.workload]); w[4] = args.dist; w[5] = args.param
        w[7] = w[7].replace("Zipf-skewed", f"{args.dist}" + (f"({args.param})" if args.param else "")).replace("uniform", args.dist)
        WORKLOADS[args.workload] = tuple(w)

    if args.impl == "reference":
        if rank != 0:
            return 0
        if torch.cuda.is_available():
            torch.cuda.set_device(0)
        print(json.dumps(run_reference(args, WORKLOADS[args.workload])), flush=True)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        line = run_ours_multi(args, rank, world)
        dist.barrier()
        dist.destroy_process_group()
    else:
        line = run_ours_single(args, WORKLOADS[args.workload])
        if args.workload == "cfg2" and not args.logn and not args.no_extras:
            line["other_configs"] = extra_workloads(args)
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
