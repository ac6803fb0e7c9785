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
