"""CPU tests of the multi-GPU path's host-side logic (gpu_sort_b200/dist.py) under the `gloo` backend, world_size 2 and 3:
histogram all-reduce, deterministic splitter selection, count all-gather, source-rank-ordered all-to-all.  The three device
operations (histogram, stable range partition, local sort) are replaced by the numpy test double below -- it lives in
tests/ only; the product's default `CudaOps` has no CPU path.  The expected result is ONE stable sort of the concatenated
input (SURVEY.md section 8e "Validation")."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class NumpyOps:
    """Test double for CudaOps: same contracts as b200_msd_histogram / b200_range_partition / the local sort."""

    def histogram(self, keys, bits):
        k = keys.numpy().view(np.uint32)
        return torch.from_numpy(np.bincount((k >> np.uint32(32 - bits)).astype(np.int64), minlength=1 << bits).astype(np.int64))

    def partition(self, keys, vals, bits, splitters, local_counts):
        k = keys.numpy().view(np.uint32)
        bucket = (k >> np.uint32(32 - bits)).astype(np.int64)
        dest = np.searchsorted(np.asarray(splitters, dtype=np.int64), bucket, side="right") if len(splitters) else np.zeros(k.size, dtype=np.int64)
        order = np.argsort(dest, kind="stable")
        offs = np.concatenate([[0], np.cumsum(np.bincount(dest, minlength=len(splitters) + 1))]).astype(np.int64)
        return keys[torch.from_numpy(order)], (vals[torch.from_numpy(order)] if vals is not None else None), torch.from_numpy(offs)

    def bucket_hist(self, keys, xbits, key_bits=32):
        k = keys.numpy().view(np.uint32)
        b = (k >> np.uint32(32 - xbits)).astype(np.int64) if xbits else np.zeros(k.size, dtype=np.int64)
        return torch.from_numpy(np.bincount(b, minlength=256).astype(np.int64))

    def bucket_partition(self, keys, vals, xbits, bound, key_bits=32):
        k = keys.numpy().view(np.uint32)
        b = (k >> np.uint32(32 - xbits)).astype(np.int64) if xbits else np.zeros(k.size, dtype=np.int64)
        order = torch.from_numpy(np.argsort(b, kind="stable"))            # by bucket = by destination then bucket (destinations own bucket ranges)
        cnt = np.bincount(b, minlength=256)
        send = [int(cnt[bound[j]:bound[j + 1]].sum()) for j in range(len(bound) - 1)]
        return keys[order], (vals[order] if vals is not None else None), send

    def local_sort(self, keys, vals, n, stable):
        order = torch.from_numpy(np.argsort(keys.numpy().view(np.uint32), kind="stable"))
        return keys[order], (vals[order] if vals is not None else None)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gen(total, dist_name, seed=7):
    rng = np.random.default_rng(seed)
    if dist_name == "uniform":
        k = rng.integers(0, 2**32, size=total, dtype=np.uint32)
    elif dist_name == "lowent":
        k = rng.integers(0, 2**32, size=total, dtype=np.uint32) & rng.integers(0, 2**32, size=total, dtype=np.uint32) & rng.integers(0, 2**32, size=total, dtype=np.uint32)
    elif dist_name == "constant":
        k = np.full(total, 0xDEADBEEF, dtype=np.uint32)
    else:
        k = np.sort(rng.integers(0, 2**32, size=total, dtype=np.uint32))
    return k


def _worker(rank, world, port, total, dist_name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpu_sort_b200 import dist as gd
        k = _gen(total, dist_name)
        lo, hi = rank * total // world, (rank + 1) * total // world
        keys = torch.from_numpy(k[lo:hi].view(np.int32).copy())
        vals = torch.arange(lo, hi, dtype=torch.int32)          # value = global index (config 5 convention)
        sk, sv, info = gd.distributed_sort(keys, vals, bits=10, ops=NumpyOps())
        q.put((rank, sk.numpy().view(np.uint32).copy(), sv.numpy().view(np.uint32).copy(), info["count"], info["splitters"], info["imbalance"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dist_name", [(2, "uniform"), (2, "lowent"), (3, "uniform"), (2, "constant"), (3, "sorted")])
def test_distributed_sort_equals_one_stable_sort(world, dist_name):
    total = 60000 + 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, dist_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k = _gen(total, dist_name)
    order = np.argsort(k, kind="stable")
    got_k = np.concatenate([r[1] for r in res]); got_v = np.concatenate([r[2] for r in res])
    assert sum(r[3] for r in res) == total
    assert np.array_equal(got_k, k[order])
    assert np.array_equal(got_v, order.astype(np.uint32))         # globally stable
    assert all(r[4] == res[0][4] for r in res)                    # every rank chose the same splitters
    if dist_name == "uniform":
        assert res[0][5] < 1.05                                   # balanced within the bucket granularity


def test_choose_splitters_properties():
    from gpu_sort_b200 import dist as gd
    rng = np.random.default_rng(1)
    for parts in (1, 2, 4, 8):
        c = rng.integers(0, 1000, size=4096).astype(np.uint64)
        s = gd.choose_splitters(c, parts)
        assert len(s) == parts - 1 and s == sorted(s) and all(0 <= b <= 4096 for b in s)
        cum = np.concatenate([[0], np.cumsum(c)])
        for j, b in enumerate(s, 1):       # no other boundary is closer to the ideal split point
            target = cum[-1] * j / parts
            assert abs(cum[b] - target) <= np.abs(cum - target).min() + 1e-9
        assert gd.choose_splitters_tensor(torch.from_numpy(c.astype(np.int64)), parts).tolist() == s
        if parts > 1:
            pc = gd.part_counts(torch.from_numpy(c.astype(np.int64)), torch.tensor(s, dtype=torch.int32), parts)
            edges = [0] + s + [4096]
            assert pc.tolist() == [int(c[edges[i]:edges[i + 1]].sum()) for i in range(parts)]
    # one bucket holds everything (constant keys): all splitters collapse around it, nothing is lost
    c = np.zeros(256, dtype=np.uint64); c[77] = 10**6
    s = gd.choose_splitters(c, 4)
    assert all(b in (77, 78) for b in s)
    m = np.array([[5, 0], [7, 0]])
    assert gd.receive_layout(m, 0) == ([5, 0], [5, 7], 12) and gd.receive_layout(m, 1) == ([7, 0], [0, 0], 0)
    assert gd.imbalance(m) == 2.0


def _worker_exchange(rank, world, port, total, dist_name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpu_sort_b200 import dist as gd
        k = _gen(total, dist_name)
        lo, hi = rank * total // world, (rank + 1) * total // world
        keys = torch.from_numpy(k[lo:hi].view(np.int32).copy())
        vals = torch.arange(lo, hi, dtype=torch.int32)
        sk, sv, info = gd.exchange_sort(keys, vals, ops=NumpyOps(), xbits=6)
        q.put((rank, sk.numpy().view(np.uint32).copy(), sv.numpy().view(np.uint32).copy(), info["count"], info["bound"], info["imbalance"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dist_name", [(2, "uniform"), (3, "uniform"), (2, "lowent"), (2, "constant"), (3, "sorted")])
def test_exchange_as_level_0_equals_one_stable_sort(world, dist_name):
    """The default multi-GPU plan (buckets dealt to ranks in contiguous balanced groups, bucket-major receive layout in source-rank
    order, segmented finish): same host logic as ExchangeSorter, collectives under gloo, device operations by the numpy double."""
    total = 50000 + 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_exchange, args=(r, world, port, total, dist_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    k = _gen(total, dist_name)
    order = np.argsort(k, kind="stable")
    got_k = np.concatenate([o[1] for o in out]); got_v = np.concatenate([o[2] for o in out])
    assert np.array_equal(got_k, k[order])
    assert np.array_equal(got_v, order.astype(np.uint32))          # values = global indices: identical to ONE stable sort
    assert sum(o[3] for o in out) == total
    assert all(o[4] == out[0][4] for o in out)                     # every rank derived the same plan
    if dist_name == "uniform":
        assert out[0][5] < 1.2


def test_exchange_plan_properties():
    from gpu_sort_b200 import dist as gd
    rng = np.random.default_rng(3)
    for G in (1, 2, 3, 8):
        m = np.zeros((G, 256), dtype=np.int64)
        m[:, :64] = rng.integers(0, 1000, size=(G, 64))
        bound, start, ok, recv = gd.exchange_plan(m, cap=None)
        assert bound[0] == 0 and bound[-1] == 256 and all(bound[i] <= bound[i + 1] for i in range(G))
        assert sum(recv) == int(m.sum())
        for j in range(G):            # inside a destination the buckets lie back to back in ascending order
            run = 0
            for d in range(bound[j], bound[j + 1]):
                assert int(start[d]) == run
                run += int(m[:, d].sum())
            assert run == recv[j]
    m = np.zeros((2, 256), dtype=np.int64); m[:, 5] = 1000          # one heavy bucket: cannot be balanced, reported through ok
    bound, start, ok, recv = gd.exchange_plan(m, cap=1200)
    assert not ok and max(recv) == 2000
