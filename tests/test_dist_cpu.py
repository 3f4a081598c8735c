"""World-size-2 gloo tests (CPU) of the document-partition host logic: the load-time exchange
of collection statistics and the cross-shard top-k merge order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeInfo:
    def __init__(self, n_docs, avg):
        self.n_docs, self.avg_doc_len = n_docs, avg


class FakeEngine:
    """Stands in for GpuVacuumEngine on the CPU: records what the exchange hands to the device."""
    device = 0

    def __init__(self, n_docs, avg, ranks, dfs):
        self._info, self._ranks, self._dfs = FakeInfo(n_docs, avg), ranks, dfs
        self.got = None

    def info(self):
        return self._info

    def local_stats(self, want_ranks=False):
        return self._dfs, (self._ranks if want_ranks else None)

    def term_at(self, i):
        return f"t{self._ranks[i]}", int(self._dfs[i])

    def set_global_stats(self, doc_base, n, avg, df):
        self.got = (doc_base, n, avg, np.array(df))


def _worker(rank, world, port, keys, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wiser_b200.dist import ShardedSearch, merge_topk_host
    from wiser_b200.capi import HIT_DTYPE
    if rank == 0:
        eng = FakeEngine(1000, 80.0, np.array([0, 1, 5], np.uint32), np.array([900, 500, 3], np.uint32))
    else:
        eng = FakeEngine(3000, 100.0, np.array([1, 0, 7, 5], np.uint32), np.array([1500, 2800, 1, 9], np.uint32))
    sh = ShardedSearch(eng, rank, world, device=torch.device("cpu"), term_keys=keys)
    # per-shard top-2 of one query, merged on every rank after an all_gather over gloo
    mine = np.zeros((1, 2), HIT_DTYPE)
    mine["doc_id"][0] = [10, 11] if rank == 0 else [1000, 1001]
    mine["score"][0] = [2.0, 1.0] if rank == 0 else [2.0, 1.5]
    t = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy())
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    g = np.stack([o.numpy().view(HIT_DTYPE).reshape(1, 2) for o in outs])
    merged, mn = merge_topk_host(g, np.full((world, 1), 2, np.int32), 3)
    q.put((rank, eng.got, merged["doc_id"][0].tolist(), merged["score"][0].tolist(), int(mn[0])))
    dist.destroy_process_group()


@pytest.mark.parametrize("keys", ["synthetic_rank", "strings"])
def test_stats_exchange_and_merge_world2(keys):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if keys == "strings" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, keys, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, got0, docs0, sc0, n0), (r1, got1, docs1, sc1, n1) = res
    # doc bases are the exclusive prefix of partition sizes; N and the average are global
    assert got0[0] == 0 and got1[0] == 1000 and got0[1] == got1[1] == 4000
    assert got0[2] == got1[2] == (80.0 * 1000 + 100.0 * 3000) / 4000
    # global df per LOCAL term id
    assert got0[3].tolist() == [900 + 2800, 500 + 1500, 3 + 9]
    assert got1[3].tolist() == [1500 + 500, 2800 + 900, 1, 9 + 3]
    # merge order: score desc, doc id asc; identical on both ranks
    assert docs0 == docs1 == [10, 1000, 1001] and sc0 == sc1 == [2.0, 2.0, 1.5] and n0 == n1 == 3


def test_combine_partition_stats_is_order_stable():
    from wiser_b200.dist import combine_partition_stats
    total, bases, avg = combine_partition_stats([5, 7, 9], [10.0, 20.0, 30.0])
    assert total == 21 and bases == [0, 5, 12]
    assert avg == ((10.0 * 5 + 20.0 * 7) + 30.0 * 9) / 21


def _scatter_worker(rank, world, port, n, k, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wiser_b200.dist import exchange_slices, merge_topk_host, slice_bounds
    from wiser_b200.capi import HIT_DTYPE
    # every shard's top-k of n queries (seeded per shard; doc ids disjoint between shards)
    rng = np.random.default_rng(100 + rank)
    hits = np.zeros((n, k), HIT_DTYPE)
    nh = rng.integers(0, k + 1, n).astype(np.int32)
    for i in range(n):
        sc = np.sort(rng.integers(1, 50, nh[i]).astype(np.float64))[::-1]     # many ties across shards
        hits["score"][i, :nh[i]] = sc
        hits["doc_id"][i, :nh[i]] = rank * 100000 + np.sort(rng.choice(1000, nh[i], replace=False))
    s, lo = slice_bounds(n, world)
    recv_h = torch.zeros(world * max(1, s) * k * 16, dtype=torch.uint8)
    recv_n = torch.zeros(world * max(1, s), dtype=torch.int32)
    mine = exchange_slices(torch.from_numpy(hits.view(np.uint8).reshape(-1).copy()), torch.from_numpy(nh.copy()),
                           n, k, rank, world, recv_h, recv_n)
    assert mine == lo[rank + 1] - lo[rank]
    g = recv_h[:world * mine * k * 16].numpy().view(HIT_DTYPE).reshape(world, mine, k)
    gn = recv_n[:world * mine].numpy().reshape(world, mine)
    m_h, m_n = merge_topk_host(g, gn, k)
    # all-gather of the merged (padded) slices: the first n queries of the result are in order
    pad_h = np.zeros((max(1, s), k), HIT_DTYPE)
    pad_n = np.zeros(max(1, s), np.int32)
    pad_h[:mine], pad_n[:mine] = m_h, m_n
    full_h = [torch.zeros(max(1, s) * k * 16, dtype=torch.uint8) for _ in range(world)]
    full_n = [torch.zeros(max(1, s), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(full_h, torch.from_numpy(pad_h.view(np.uint8).reshape(-1).copy()))
    dist.all_gather(full_n, torch.from_numpy(pad_n.copy()))
    res_h = torch.cat(full_h).numpy().view(HIT_DTYPE).reshape(-1, k)[:n]
    res_n = torch.cat(full_n).numpy()[:n]
    # checker: the plain all-gather of everything, merged in one piece
    all_h = [torch.zeros(n * k * 16, dtype=torch.uint8) for _ in range(world)]
    all_n = [torch.zeros(n, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(all_h, torch.from_numpy(hits.view(np.uint8).reshape(-1).copy()))
    dist.all_gather(all_n, torch.from_numpy(nh.copy()))
    ref_h, ref_n = merge_topk_host(np.stack([t.numpy().view(HIT_DTYPE).reshape(n, k) for t in all_h]),
                                   np.stack([t.numpy() for t in all_n]), k)
    ok = bool(np.array_equal(res_n, ref_n))
    for i in range(n):
        ok = ok and np.array_equal(res_h[i, :ref_n[i]], ref_h[i, :ref_n[i]])
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(7, 3), (8, 10), (1, 2), (33, 1)])
def test_scatter_exchange_equals_allgather_merge_world2(n, k):
    """The scatter exchange (all-to-all of query slices -> slice merge -> all-gather of merged
    slices) returns what merging a plain all-gather returns, for ragged and tiny batches."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + 17 * n + k) % 300
    ps = [ctx.Process(target=_scatter_worker, args=(r, 2, port, n, k, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert got == [(0, True), (1, True)]


def test_slice_bounds_cover_every_query_once():
    from wiser_b200.dist import slice_bounds
    for n in (0, 1, 7, 8, 9, 100000, 100003):
        for world in (1, 2, 3, 8):
            s, lo = slice_bounds(n, world)
            assert lo[0] == 0 and lo[-1] == n and len(lo) == world + 1
            sizes = [lo[r + 1] - lo[r] for r in range(world)]
            assert all(0 <= x <= s for x in sizes) and sum(sizes) == n
            # slice r starts at r*s: the all-gather of s-padded slices is the result in query order
            assert all(lo[r] == min(n, r * s) for r in range(world))


def test_sharded_search_refuses_doc_range_shards():
    """A doc-range shard of ONE directory (wsr_index_open(dir, dev, shard, n_shards)) already scores
    with the collection's statistics and emits global doc ids: the partition exchange would count N
    world times and offset ids twice, so ShardedSearch refuses such an engine."""
    sys.path.insert(0, ROOT)
    from wiser_b200.dist import ShardedSearch
    eng = FakeEngine(1000, 80.0, np.array([0], np.uint32), np.array([1], np.uint32))
    eng.n_shards = 2
    with pytest.raises(ValueError, match="OWN partition directory"):
        ShardedSearch(eng, 0, 2, device=torch.device("cpu"))
