"""gpu_sort_b200.dist -- the multi-GPU sort (BASELINE.json config 5).

The reference is single-GPU (no NCCL/MPI/P2P call anywhere in /root/reference); this is the one place where the
sort shards naturally, with exactly ONE exchange step (a range-partition / sample sort):

  1. local histogram of the top `bits` bits of the order-transformed keys         b200_msd_histogram      (CUDA)
  2. all-reduce(sum) of the histograms -> every rank sees the global distribution  torch.distributed / NCCL
  3. every rank picks the same G-1 bucket splitters (cumulative count closest to j*n/G)   choose_splitters
  4. stable G-way split of the local (key, value) pairs into contiguous send segments     b200_range_partition (CUDA)
  5. all-gather of the G send counts -> GxG matrix -> receive offsets in SOURCE-RANK order
  6. key/value all-to-all over NVLink (grouped send/recv inside NCCL's all_to_all_single)
  7. independent local sort of what was received: stable LSB sort for pairs (steps 4+6+7 stable => the global
     result equals ONE stable sort of the concatenated input), MSB hybrid sort for keys-only.

One process per GPU; `torch.distributed` is plumbing only.  The device work of steps 1, 4 and 7 goes through the C ABI
(include/b200sort.h).  `ops` is the seam the CPU (gloo) tests use to exercise the host-side logic of steps 2, 3, 5, 6
without a GPU; the default `CudaOps` has no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

DEFAULT_BITS = 14      # 16384 buckets: splitters are fine-grained (n/G/2048 per bucket at G=8 on uniform keys)


# ----------------------------------------------------------------------------------------------------------------
# device operations (C ABI)
# ----------------------------------------------------------------------------------------------------------------
def msd_histogram(keys: torch.Tensor, bits: int = DEFAULT_BITS, key_type: Optional[int] = None, stream=None) -> torch.Tensor:
    """counts[b] = #keys whose top `bits` bits (order-transformed) equal b; int64[1 << bits] on the keys' device."""
    import gpu_sort_b200 as gs
    counts = torch.empty(1 << bits, dtype=torch.int64, device=keys.device)
    gs._check(gs.lib.b200_msd_histogram(gs._ptr(keys), keys.numel(), gs.key_type_of(keys, key_type), bits, gs._ptr(counts), gs._stream(stream)),
              "b200_msd_histogram")
    return counts


def range_partition(keys: torch.Tensor, vals: Optional[torch.Tensor], bits: int, splitters: Sequence[int], local_counts: torch.Tensor,
                    key_type: Optional[int] = None, stream=None, out_keys=None, out_vals=None):
    """Stable split of (keys, vals) into len(splitters)+1 contiguous parts; part of a key = #{j: splitters[j] <= bucket}.
    Returns (keys_out, vals_out, part_offsets[int64, parts+1, device])."""
    import gpu_sort_b200 as gs
    parts = len(splitters) + 1
    n = keys.numel()
    ko = torch.empty_like(keys) if out_keys is None else out_keys
    vo = None if vals is None else (torch.empty_like(vals) if out_vals is None else out_vals)
    sp = torch.tensor(list(splitters) + [0], dtype=torch.int32, device=keys.device) if not isinstance(splitters, torch.Tensor) else splitters
    offs = torch.zeros(parts + 1, dtype=torch.int64, device=keys.device)
    kt = gs.key_type_of(keys, key_type)
    vb = gs._value_bytes(vals)
    nbytes = ctypes.c_size_t(0)
    args = lambda temp: (gs._ptr(temp), ctypes.byref(nbytes), gs._ptr(keys), gs._ptr(vals), gs._ptr(ko), gs._ptr(vo), n, kt, vb, bits,
                         gs._ptr(sp), parts, gs._ptr(local_counts), gs._ptr(offs), gs._stream(stream))
    gs._check(gs.lib.b200_range_partition(*args(None)), "b200_range_partition(size query)")
    temp = torch.empty(nbytes.value, dtype=torch.uint8, device=keys.device)
    gs._check(gs.lib.b200_range_partition(*args(temp)), "b200_range_partition")
    return ko, vo, offs


class CudaOps:
    """Steps 1, 4 and 7 on the GPU through the C ABI.  There is no CPU implementation in the product."""

    def __init__(self, key_type: Optional[int] = None):
        import gpu_sort_b200 as gs     # raises ImportError when libb200sort.so is missing
        self.gs = gs
        self.key_type = key_type

    def histogram(self, keys, bits):
        return msd_histogram(keys, bits, self.key_type)

    def partition(self, keys, vals, bits, splitters, local_counts):
        return range_partition(keys, vals, bits, splitters, local_counts, self.key_type)

    def local_sort(self, keys, vals, n, stable):
        gs = self.gs
        kt = gs.key_type_of(keys, self.key_type)
        k_alt = torch.empty_like(keys)
        v_alt = torch.empty_like(vals) if vals is not None else None
        if stable or vals is not None:
            dk = gs.DoubleBuffer(keys, k_alt)
            dv = gs.DoubleBuffer(vals, v_alt) if vals is not None else None
            tb = gs.DeviceRadixSort._run(None, dk, dv, n, 0, None, False, None, kt)
            temp = torch.empty(tb, dtype=torch.uint8, device=keys.device)
            gs.DeviceRadixSort._run(temp, dk, dv, n, 0, None, False, None, kt)
            return dk.Current(), (dv.Current() if vals is not None else None)
        r = gs.rdxsrt_unstable_sort(keys, None, n, k_alt, None, key_type=kt)
        return r.sorted_keys, None


# ----------------------------------------------------------------------------------------------------------------
# host-side logic (runs identically under NCCL on GPUs and under gloo on CPU tensors)
# ----------------------------------------------------------------------------------------------------------------
def choose_splitters(global_counts, parts: int) -> List[int]:
    """G-1 ascending bucket indices; part j receives the buckets [splitter[j-1], splitter[j]).  Splitter j is the bucket
    boundary whose cumulative count is closest to j*n/G (ties -> the smaller boundary), so every rank derives the same
    splitters from the same all-reduced histogram.  A bucket heavier than n/G cannot be split by key range: the
    neighbouring parts simply come out imbalanced (reported by `imbalance`)."""
    c = np.asarray(global_counts.cpu() if isinstance(global_counts, torch.Tensor) else global_counts).astype(np.uint64, copy=False)
    nb = c.size
    cum = np.concatenate([[0], np.cumsum(c, dtype=np.uint64)]).astype(np.float64)      # cum[b] = #keys in buckets < b
    n = cum[-1]
    out: List[int] = []
    for j in range(1, parts):
        target = n * j / parts
        b = int(np.searchsorted(cum, target, side="left"))          # first boundary with cum >= target
        b = min(max(b, 0), nb)
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[b] - target):
            b -= 1
        if out and b < out[-1]:
            b = out[-1]
        out.append(b)
    return out


def choose_splitters_tensor(global_counts: torch.Tensor, parts: int) -> torch.Tensor:
    """choose_splitters with torch ops on the tensor's own device (no device->host copy of the histogram); int32[parts-1].
    Same rule, same tie-break: tests/test_dist_cpu.py checks it against the host version."""
    c = global_counts.to(torch.float64)
    cum = torch.cat([torch.zeros(1, dtype=torch.float64, device=c.device), torch.cumsum(c, 0)])
    n = cum[-1]
    targets = n * torch.arange(1, parts, dtype=torch.float64, device=c.device) / parts
    b = torch.searchsorted(cum, targets, right=False).clamp_(0, c.numel())
    lower = (b - 1).clamp_(min=0)
    take_lower = (b > 0) & ((cum[lower] - targets).abs() <= (cum[b] - targets).abs())
    b = torch.where(take_lower, lower, b)
    b = torch.cummax(b, 0).values if parts > 1 else b
    return b.to(torch.int32)


def part_counts(local_counts: torch.Tensor, splitters: torch.Tensor, parts: int) -> torch.Tensor:
    """Keys this rank sends to every destination: sum of its own bucket counts inside each splitter range; int64[parts]."""
    bucket = torch.arange(local_counts.numel(), device=local_counts.device, dtype=torch.int32)
    dest = torch.searchsorted(splitters.to(torch.int32), bucket, right=True) if parts > 1 else torch.zeros_like(bucket, dtype=torch.int64)
    out = torch.zeros(parts, dtype=torch.int64, device=local_counts.device)
    out.scatter_add_(0, dest.to(torch.int64), local_counts.to(torch.int64))
    return out


def receive_layout(count_matrix: np.ndarray, rank: int) -> Tuple[List[int], List[int], int]:
    """count_matrix[src][dst] = keys src sends to dst.  Returns (send_sizes, recv_sizes, n_recv) for `rank`; the receive
    buffer is filled in source-rank order, which is what keeps the distributed sort stable."""
    send = [int(x) for x in count_matrix[rank, :]]
    recv = [int(x) for x in count_matrix[:, rank]]
    return send, recv, int(sum(recv))


def imbalance(count_matrix: np.ndarray) -> float:
    """max over ranks of (keys received) / (n / G)."""
    per = count_matrix.sum(axis=0).astype(np.float64)
    mean = per.sum() / max(len(per), 1)
    return float(per.max() / mean) if mean > 0 else 1.0


def distributed_sort(keys: torch.Tensor, vals: Optional[torch.Tensor] = None, group=None, bits: int = DEFAULT_BITS, ops=None,
                     stable: bool = True, key_type: Optional[int] = None, timings: Optional[dict] = None):
    """Sorts the union of every rank's (keys, vals).  Rank r ends up with the r-th key range, sorted; returns
    (sorted_keys, sorted_vals, info) with info = {"count", "count_matrix", "splitters", "imbalance"}.
    Input buffers are clobbered (like the single-GPU entry points)."""
    if ops is None:
        ops = CudaOps(key_type)
    G = dist.get_world_size(group)
    rank = dist.get_rank(group)

    local_counts = ops.histogram(keys, bits)                                  # 1
    global_counts = local_counts
    if G > 1:
        global_counts = local_counts.clone()                                  # the partition needs the LOCAL counts later
        dist.all_reduce(global_counts, op=dist.ReduceOp.SUM, group=group)     # 2  (the "global MSB histogram allreduce")
    splitters = choose_splitters(global_counts, G)                            # 3  (one small D2H: 8 << bits bytes)
    pk, pv, offs = ops.partition(keys, vals, bits, splitters, local_counts)   # 4

    send_counts = (offs[1:] - offs[:-1]).contiguous()                         # 5
    gathered = [torch.empty_like(send_counts) for _ in range(G)]
    if G > 1:
        dist.all_gather(gathered, send_counts, group=group)
    else:
        gathered = [send_counts]
    matrix = np.stack([g.cpu().numpy() for g in gathered]).astype(np.int64)   # [src][dst]
    send, recv, n_recv = receive_layout(matrix, rank)

    rk = torch.empty(max(n_recv, 1), dtype=keys.dtype, device=keys.device)[:n_recv]      # 6
    rv = torch.empty(max(n_recv, 1), dtype=vals.dtype, device=vals.device)[:n_recv] if vals is not None else None
    if G > 1:
        dist.all_to_all_single(rk, pk, output_split_sizes=recv, input_split_sizes=send, group=group)
        if vals is not None:
            dist.all_to_all_single(rv, pv, output_split_sizes=recv, input_split_sizes=send, group=group)
    else:
        rk.copy_(pk)
        if vals is not None:
            rv.copy_(pv)

    sk, sv = ops.local_sort(rk, rv, n_recv, stable) if n_recv else (rk, rv)   # 7
    info = {"count": n_recv, "count_matrix": matrix, "splitters": splitters, "imbalance": imbalance(matrix)}
    return sk, sv, info


# ----------------------------------------------------------------------------------------------------------------
# The CUDA fast path: buffers allocated once, exchange fused into the scatter kernel over NVLink peer memory
# ----------------------------------------------------------------------------------------------------------------
class DistSorter:
    """Multi-GPU sort with everything allocated up front (one process per GPU).

    fused=True (default): the receive buffers live in symmetric memory (torch.distributed._symmetric_memory: every rank maps
    every peer's buffer); the range-partition kernel writes each destination's part STRAIGHT into that GPU's receive buffer
    (b200_range_partition_to: coalesced 128-byte runs over NVLink from inside the scatter's write-out), so the key/value
    all-to-all is not a separate collective and overlaps the partition tile by tile.  The only collectives left are the
    512 KB histogram all-reduce, the GxG count all-gather and two barriers.
    fused=False: the NCCL baseline (partition into a local send buffer, then all_to_all_single).
    """

    def __init__(self, n_local: int, key_dtype=torch.int32, value_dtype=None, group=None, bits: int = DEFAULT_BITS, slack: float = 1.15,
                 fused: bool = True, key_type: Optional[int] = None, stable: Optional[bool] = None):
        import gpu_sort_b200 as gs
        self.gs, self.group, self.bits = gs, group, bits
        self.G, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.kt = key_type if key_type is not None else gs._TORCH_KEY[key_dtype]
        self.pairs = value_dtype is not None
        self.stable = self.pairs if stable is None else stable
        self.n_local = n_local
        self.cap = int(n_local * slack) + 4096
        dev = torch.device("cuda", torch.cuda.current_device())
        self.fused = fused and self.G > 1
        self.hk = self.hv = None
        if self.fused:
            try:
                import torch.distributed._symmetric_memory as symm
                gname = (group or dist.group.WORLD).group_name
                self.recv_k = symm.empty(self.cap, dtype=key_dtype, device=dev)
                self.hk = symm.rendezvous(self.recv_k, gname)
                ptrs_k = [int(p) for p in self.hk.buffer_ptrs]
                ptrs_v = [0] * self.G
                if self.pairs:
                    self.recv_v = symm.empty(self.cap, dtype=value_dtype, device=dev)
                    self.hv = symm.rendezvous(self.recv_v, gname)
                    ptrs_v = [int(p) for p in self.hv.buffer_ptrs]
                self.dst_k = torch.tensor(ptrs_k, dtype=torch.int64, device=dev)
                self.dst_v = torch.tensor(ptrs_v, dtype=torch.int64, device=dev)
            except Exception as e:       # no peer mapping on this system: NCCL exchange (still the CUDA path, nothing on the CPU)
                self.fused = False
                self.fused_error = repr(e)
        if not self.fused:
            self.recv_k = torch.empty(self.cap, dtype=key_dtype, device=dev)
            self.recv_v = torch.empty(self.cap, dtype=value_dtype, device=dev) if self.pairs else None
            self.send_k = torch.empty(n_local, dtype=key_dtype, device=dev)
            self.send_v = torch.empty(n_local, dtype=value_dtype, device=dev) if self.pairs else None
        elif not self.pairs:
            self.recv_v = None
        self.alt_k = torch.empty(self.cap, dtype=key_dtype, device=dev)
        self.alt_v = torch.empty(self.cap, dtype=value_dtype, device=dev) if self.pairs else None
        self.counts = torch.empty(1 << bits, dtype=torch.int64, device=dev)
        self.offs = torch.zeros(self.G + 1, dtype=torch.int64, device=dev)
        vb = gs._value_bytes(self.recv_v)
        self.vb = vb
        nb = ctypes.c_size_t(0)
        gs._check(gs.lib.b200_range_partition(None, ctypes.byref(nb), None, None, None, None, n_local, self.kt, vb, bits, None, self.G, None, None, None),
                  "b200_range_partition(size query)")
        self.part_temp = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        if self.stable:
            tb = ctypes.c_size_t(0)
            gs._check(gs.lib.b200_lsb_sort(None, ctypes.byref(tb), None, None, None, None, None, self.cap, self.kt, vb, 0, gs.KEY_BYTES[self.kt] * 8, 0, 1, None),
                      "b200_lsb_sort(size query)")
            self.sort_temp = torch.empty(tb.value, dtype=torch.uint8, device=dev)
        else:
            self.sort_temp = torch.empty(gs.rdxsrt_workspace_bytes(self.cap, self.kt, vb), dtype=torch.uint8, device=dev)

    def sort(self, keys: torch.Tensor, vals: Optional[torch.Tensor] = None, profile: bool = False):
        gs, G, rank, bits = self.gs, self.G, self.rank, self.bits
        n = keys.numel()
        stream = gs._stream(None)
        marks = []

        def mark(name):
            if profile:
                e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
        if profile:
            gs.prof_enable(True)
        mark("start")
        gs._check(gs.lib.b200_msd_histogram(gs._ptr(keys), n, self.kt, bits, gs._ptr(self.counts), stream), "b200_msd_histogram")       # 1
        glob = self.counts
        if G > 1:
            glob = self.counts.clone()
            mark("histogram")
            dist.all_reduce(glob, group=self.group)                                                                                   # 2
        mark("allreduce")
        sp = choose_splitters_tensor(glob, G)                                                                                         # 3 (device)
        spp = torch.cat([sp, sp.new_zeros(1)])
        mine = part_counts(self.counts, sp, G)                                                                                        # 5
        if G > 1:
            mat = torch.empty(G * G, dtype=torch.int64, device=keys.device)
            dist.all_gather_into_tensor(mat, mine, group=self.group)
            mat = mat.view(G, G)
        else:
            mat = mine.view(1, 1)
        base = mat[:rank].sum(0) if rank > 0 else torch.zeros(G, dtype=torch.int64, device=keys.device)      # my start inside every destination
        mark("splitters+counts")
        host = torch.cat([mat.reshape(-1), sp.to(torch.int64)]).cpu().numpy()     # the one host sync: n_recv sizes the local sort
        matrix = host[:G * G].reshape(G, G)
        send, recv, n_recv = receive_layout(matrix, rank)
        # every key this rank receives lies in [lo, hi]: the leading bits on which lo and hi agree need not be sorted
        kbits = gs.KEY_BYTES[self.kt] * 8
        lo = (int(host[G * G + rank - 1]) if rank > 0 else 0) << (kbits - bits)
        hi = (((int(host[G * G + rank]) if rank < G - 1 else (1 << bits)) << (kbits - bits)) - 1) if G > 1 else (1 << kbits) - 1
        end_bit = max((lo ^ max(hi, lo)).bit_length(), 1)
        if n_recv > self.cap:
            raise RuntimeError(f"rank {rank}: {n_recv} keys to receive exceed the receive capacity {self.cap} (key range too skewed for range partitioning)")
        nb = ctypes.c_size_t(self.part_temp.numel())
        if self.fused:
            self.hk.barrier()                                        # every peer is done with the previous contents of its receive buffer
            mark("d2h+barrier")
            gs._check(gs.lib.b200_range_partition_to(gs._ptr(self.part_temp), ctypes.byref(nb), gs._ptr(keys), gs._ptr(vals), n, self.kt, self.vb, bits,
                                                     gs._ptr(spp), G, gs._ptr(self.counts), gs._ptr(self.offs), gs._ptr(self.dst_k), gs._ptr(self.dst_v),
                                                     gs._ptr(base.contiguous()), stream), "b200_range_partition_to")                  # 4 + 6 fused
            mark("partition_fused")
            self.hk.barrier()                                        # all peers' stores have landed
            mark("barrier")
        else:
            gs._check(gs.lib.b200_range_partition(gs._ptr(self.part_temp), ctypes.byref(nb), gs._ptr(keys), gs._ptr(vals), gs._ptr(self.send_k),
                                                  gs._ptr(self.send_v), n, self.kt, self.vb, bits, gs._ptr(spp), G, gs._ptr(self.counts), gs._ptr(self.offs),
                                                  stream), "b200_range_partition")                                                    # 4
            mark("partition")
            if G > 1:
                dist.all_to_all_single(self.recv_k[:n_recv], self.send_k, output_split_sizes=recv, input_split_sizes=send, group=self.group)   # 6
                if self.pairs:
                    dist.all_to_all_single(self.recv_v[:n_recv], self.send_v, output_split_sizes=recv, input_split_sizes=send, group=self.group)
            else:
                self.recv_k[:n].copy_(self.send_k)
                if self.pairs:
                    self.recv_v[:n].copy_(self.send_v)
        mark("exchange")
        rk = self.recv_k[:n_recv]; rv = self.recv_v[:n_recv] if self.pairs else None
        if n_recv == 0:
            sk, sv = rk, rv
        elif self.stable:                                                                                                             # 7
            dk = gs.DoubleBuffer(rk, self.alt_k[:n_recv]); dv = gs.DoubleBuffer(rv, self.alt_v[:n_recv]) if self.pairs else None
            gs.DeviceRadixSort._run(self.sort_temp, dk, dv, n_recv, 0, end_bit, False, None, self.kt)
            sk, sv = dk.Current(), (dv.Current() if self.pairs else None)
        else:
            r = gs.rdxsrt_unstable_sort(rk, rv, n_recv, self.alt_k[:n_recv], self.alt_v[:n_recv] if self.pairs else None, workspace=self.sort_temp, key_type=self.kt,
                                        end_bit=end_bit)
            sk, sv = r.sorted_keys, r.sorted_values
        mark("local_sort")
        info = {"count": n_recv, "count_matrix": matrix, "imbalance": imbalance(matrix), "fused": self.fused}
        if profile:
            torch.cuda.synchronize()
            info["kernels_ms"] = {k: round(v[1], 3) for k, v in gs.prof_report().items()}
            gs.prof_enable(False)
            info["phases_ms"] = {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 3) for i in range(1, len(marks))}
        return sk, sv, info
