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
        self.key_dtype, self.value_dtype = key_dtype, value_dtype
        self.fused = fused and self.G > 1
        self._alloc(int(n_local * slack) + 4096)

    def _alloc(self, cap: int):
        """(Re)allocates everything sized by the receive capacity.  Collective when fused (symmetric-memory rendezvous): every
        rank calls it with the same capacity -- sort() derives it from the all-gathered count matrix, identical on all ranks."""
        gs, group, n_local, bits = self.gs, self.group, self.n_local, self.bits
        key_dtype, value_dtype = self.key_dtype, self.value_dtype
        self.cap = cap
        dev = torch.device("cuda", torch.cuda.current_device())
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
        if not hasattr(self, "counts"):          # (kept across a re-allocation: sort() grows the buffers after the histogram is taken)
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
        worst = int(matrix.sum(axis=0).max())
        if worst > self.cap:
            # a key range heavier than the receive capacity (heavy duplicates: one bucket cannot be split by key range): tolerated as
            # imbalance (reported in info) -- every rank sees the same matrix, so all of them grow their buffers together
            self._alloc(int(worst * 1.02) + 4096)
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


# ----------------------------------------------------------------------------------------------------------------
# The default multi-GPU path: the exchange IS level 0 of the sort
# ----------------------------------------------------------------------------------------------------------------
def exchange_plan(matrix: np.ndarray, cap: Optional[int] = None):
    """Host restatement of exchange_plan_kernel (csrc/sort_impl.cuh), used by the CPU tests and to cross-check the device plan.
    matrix[r][d] = keys of source rank r whose leading digit is d.  Returns (bound, start, ok): rank j owns the digits
    [bound[j], bound[j+1]); start[d] = index of digit d inside its destination's receive buffer; ok = nobody exceeds `cap`."""
    m = np.asarray(matrix, dtype=np.uint64)
    G, R = m.shape
    tot = m.sum(axis=0)
    cum = np.concatenate([[0], np.cumsum(tot, dtype=np.uint64)]).astype(np.uint64)
    n = int(cum[-1])
    bound = [0]
    for j in range(1, G):
        target = (n * j) // G
        b = int(np.searchsorted(cum, np.uint64(target), side="left"))
        if b > 0 and target - int(cum[b - 1]) <= int(cum[b]) - target:
            b -= 1
        bound.append(max(b, bound[-1]))
    bound.append(R)
    start = np.zeros(R, dtype=np.uint64)
    for j in range(G):
        for d in range(bound[j], bound[j + 1]):
            start[d] = cum[d] - cum[bound[j]]
    recv = [int(cum[bound[j + 1]] - cum[bound[j]]) for j in range(G)]
    return bound, start, (cap is None or max(recv) <= cap), recv


def default_xbits(n_total: int, G: int) -> int:
    """Exchange buckets = leading xbits bits of the key: enough to deal whole buckets to the ranks (2 extra bits when G is not a
    power of two), few enough that a scatter tile's run per bucket is long (NVLink wants >= 128-byte pieces), and such that two
    8-bit levels of the local finish leave ~4096-key buckets: log2(total keys) - 16 - 12."""
    total_bits = max(int(np.ceil(np.log2(max(n_total, 2)))), 1)
    need = int(np.ceil(np.log2(G))) + (0 if (G & (G - 1)) == 0 else 2)
    return int(min(8, max(need, total_bits - 28)))


def exchange_sort(keys: torch.Tensor, vals: Optional[torch.Tensor] = None, group=None, xbits: Optional[int] = None, ops=None, key_bits: int = 32):
    """The exchange-as-level-0 sort with the collectives spelled out (all_gather of the bucket counts, all_to_all of the parts):
    the same plan ExchangeSorter executes on the GPUs with the scatter fused into the exchange, runnable under gloo with a test
    double for the three device operations (`ops.bucket_hist`, `ops.bucket_partition`, `ops.local_sort`)."""
    G = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if xbits is None:
        n_all = torch.tensor([keys.numel()], dtype=torch.int64)
        if G > 1:
            dist.all_reduce(n_all, group=group)
        xbits = default_xbits(int(n_all.item()), G)
    hist = ops.bucket_hist(keys, xbits, key_bits)                              # int64[256]
    gathered = [torch.empty_like(hist) for _ in range(G)]
    if G > 1:
        dist.all_gather(gathered, hist, group=group)
    else:
        gathered = [hist]
    matrix = np.stack([g.cpu().numpy() for g in gathered]).astype(np.int64)   # [src][bucket]
    bound, start, ok, recv_counts = exchange_plan(matrix)
    pk, pv, send = ops.bucket_partition(keys, vals, xbits, bound, key_bits)    # stable, ordered by destination then bucket
    recv = [int(matrix[r, bound[rank]:bound[rank + 1]].sum()) for r in range(G)]
    n_recv = sum(recv)
    rk = torch.empty(max(n_recv, 1), dtype=keys.dtype)[:n_recv]
    rv = torch.empty(max(n_recv, 1), dtype=vals.dtype)[:n_recv] if vals is not None else None
    if G > 1:
        dist.all_to_all_single(rk, pk, output_split_sizes=recv, input_split_sizes=send, group=group)
        if vals is not None:
            dist.all_to_all_single(rv, pv, output_split_sizes=recv, input_split_sizes=send, group=group)
    else:
        rk.copy_(pk)
        if vals is not None:
            rv.copy_(pv)
    sk, sv = ops.local_sort(rk, rv, n_recv, True) if n_recv else (rk, rv)
    return sk, sv, {"count": n_recv, "bound": bound, "xbits": xbits, "imbalance": max(recv_counts) / max(sum(recv_counts) / G, 1), "matrix": matrix}


class ExchangeSorter:
    """Multi-GPU sort of 4-/8-byte keys (+ values), one process per GPU, everything allocated up front, NO host read-back
    inside sort():

        count   : per-tile histogram of the leading 8-bit digit, one read of the keys        b200_exchange_hist
        gather  : all_gather of every rank's 256 digit totals (G x 2 KB)                      NCCL
        scatter : every rank derives the same plan on the device (digits dealt to ranks in contiguous balanced groups) and
                  runs the stable scatter of a sort level whose destinations are the PEERS' receive buffers
                  (symmetric memory, stores over NVLink from the scatter's coalesced write-out) b200_exchange_scatter
        finish  : each rank now holds level-0 buckets of the global sort: segmented sort on the remaining bits
                                                                                               b200_segmented_sort
    Stable for pairs: inside a digit the keys arrive in source-rank order, the scatter and the segmented sort are stable, so
    the global result equals ONE stable sort of the concatenated input.  The received count, the status word (1 = a rank
    would receive more than its capacity: nothing was exchanged, use DistSorter's key-range path) and the largest receive
    count live in `self.info` on the device; result() reads them back.
    """

    def __init__(self, n_local: int, key_dtype=torch.int32, value_dtype=None, group=None, slack: float = 1.15, key_type: Optional[int] = None, xbits: Optional[int] = None):
        import gpu_sort_b200 as gs
        import torch.distributed._symmetric_memory as symm
        self.gs, self.group = gs, group
        self.G, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.kt = key_type if key_type is not None else gs._TORCH_KEY[key_dtype]
        self.kbits = gs.KEY_BYTES[self.kt] * 8
        self.pairs = value_dtype is not None
        self.n_local = n_local
        self.cap = int(n_local * slack) + 4096
        dev = torch.device("cuda", torch.cuda.current_device())
        gname = (group or dist.group.WORLD).group_name
        self.recv_k = symm.empty(self.cap, dtype=key_dtype, device=dev)
        self.hk = symm.rendezvous(self.recv_k, gname)
        ptrs_k = [int(p) for p in self.hk.buffer_ptrs]
        ptrs_v = [0] * self.G
        self.recv_v = None
        if self.pairs:
            self.recv_v = symm.empty(self.cap, dtype=value_dtype, device=dev)
            self.hv = symm.rendezvous(self.recv_v, gname)
            ptrs_v = [int(p) for p in self.hv.buffer_ptrs]
        self.dst_k = torch.tensor(ptrs_k, dtype=torch.int64, device=dev)
        self.dst_v = torch.tensor(ptrs_v, dtype=torch.int64, device=dev)
        self.alt_k = torch.empty(self.cap, dtype=key_dtype, device=dev)
        self.alt_v = torch.empty(self.cap, dtype=value_dtype, device=dev) if self.pairs else None
        self.hist = torch.zeros(256, dtype=torch.int64, device=dev)
        self.matrix = torch.zeros(self.G * 256, dtype=torch.int64, device=dev)
        self.seg_begin = torch.zeros(256, dtype=torch.int64, device=dev)
        self.seg_end = torch.zeros(256, dtype=torch.int64, device=dev)
        self.info = torch.zeros(8, dtype=torch.int64, device=dev)
        self.vb = gs._value_bytes(self.recv_v)
        # exchange buckets = leading xbits bits of the key: enough of them to deal whole buckets to the ranks, few enough that a
        # scatter tile's run per bucket is long (NVLink wants >= 128-byte pieces), and such that two 8-bit levels of the local
        # finish leave buckets that fit on chip (~4096 keys): xbits = log2(total keys) - 16 - 12
        self.xbits = default_xbits(n_local * self.G, self.G) if xbits is None else int(xbits)
        self.shift = self.kbits - self.xbits
        nb = ctypes.c_size_t(0)
        gs._check(gs.lib.b200_exchange_hist(None, ctypes.byref(nb), None, n_local, self.kt, self.vb, self.xbits, None, None), "b200_exchange_hist(size query)")
        self.x_temp = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        tb = ctypes.c_size_t(0)
        gs._check(gs.lib.b200_segmented_sort(None, ctypes.byref(tb), None, None, None, None, None, self.cap, 256, None, None, 8, self.kt, self.vb, 0, self.shift, 0, 1, None),
                  "b200_segmented_sort(size query)")
        self.s_temp = torch.empty(tb.value, dtype=torch.uint8, device=dev)
        self._last = None
        self._fallback = None

    def sort(self, keys: torch.Tensor, vals: Optional[torch.Tensor] = None, profile: bool = False):
        """Enqueues the whole sort on the current stream; returns nothing the host has to wait for.  result() hands out the tensors."""
        gs, G = self.gs, self.G
        n = keys.numel()
        stream = gs._stream(None)
        marks = []

        def mark(name):
            if profile:
                e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
        if profile:
            gs.prof_enable(True)
        mark("start")
        nb = ctypes.c_size_t(self.x_temp.numel())
        gs._check(gs.lib.b200_exchange_hist(gs._ptr(self.x_temp), ctypes.byref(nb), gs._ptr(keys), n, self.kt, self.vb, self.xbits, gs._ptr(self.hist), stream),
                  "b200_exchange_hist")
        mark("count")
        if G > 1:
            dist.all_gather_into_tensor(self.matrix, self.hist, group=self.group)
        else:
            self.matrix.copy_(self.hist)
        mark("gather")
        self.hk.barrier()                                            # every peer is done with the previous contents of its receive buffer
        mark("barrier0")
        gs._check(gs.lib.b200_exchange_scatter(gs._ptr(self.x_temp), ctypes.byref(nb), gs._ptr(keys), gs._ptr(vals), n, self.kt, self.vb, self.xbits,
                                               gs._ptr(self.matrix), G, self.rank, self.cap, gs._ptr(self.dst_k), gs._ptr(self.dst_v),
                                               gs._ptr(self.seg_begin), gs._ptr(self.seg_end), gs._ptr(self.info), stream), "b200_exchange_scatter")
        mark("scatter")
        self.hk.barrier()                                            # all peers' stores have landed
        mark("barrier1")
        dk = gs.DoubleBuffer(self.recv_k, self.alt_k)
        dv = gs.DoubleBuffer(self.recv_v, self.alt_v) if self.pairs else None
        gs.DeviceSegmentedRadixSort._run(self.s_temp, dk, dv, self.cap, 256, self.seg_begin, self.seg_end, 0, self.shift, False, None, self.kt,
                                         ties_are_equal=True)          # every bucket shares its leading xbits bits
        mark("finish")
        self._last = (dk.Current(), dv.Current() if self.pairs else None, keys, vals)
        if profile:
            torch.cuda.synchronize()
            launches = gs.prof_launches()
            rep = gs.prof_report()
            self.profile = {"kernels_ms": {k: round(v[1], 3) for k, v in rep.items()}, "launches": launches,
                            "phases_ms": {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 3) for i in range(1, len(marks))}}
            gs.prof_enable(False)

    def result(self):
        """(sorted_keys, sorted_values, info) of the last sort(); synchronises (one 64-byte read-back).  If the digit-aligned plan
        could not balance the ranks within the receive capacity, the sort is redone through DistSorter's key-range path."""
        sk, sv, keys, vals = self._last
        host = self.info.cpu().numpy()
        n_recv, status, worst = int(host[0]), int(host[1]), int(host[2])
        info = {"count": n_recv, "status": status, "imbalance": worst / max(self.n_local, 1), "fused": True, "path": "exchange-as-level-0",
                "digits": (int(host[3]), int(host[4]))}
        if status != 0:
            if self._fallback is None:
                self._fallback = DistSorter(self.n_local, keys.dtype, vals.dtype if vals is not None else None, self.group, key_type=self.kt)
            k, v, finfo = self._fallback.sort(keys, vals)
            finfo["path"] = "key-range fallback"; finfo["status"] = status
            return k, v, finfo
        return sk[:n_recv], (sv[:n_recv] if sv is not None else None), info
