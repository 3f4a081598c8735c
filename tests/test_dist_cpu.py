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
