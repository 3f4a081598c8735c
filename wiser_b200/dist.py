"""Document-partitioned multi-GPU search: one process per GPU, each holding one
partition of the corpus as its own HBM-resident index; every query goes to every partition.
Per batch the shard top-k lists are exchanged over NVLink and merged by the cross-shard merge
kernel (wsr_merge_topk_device). Two exchanges are implemented:

  "scatter" (default)  an all-to-all hands rank r every shard's lists of ITS slice of the
                       queries, rank r merges that slice, and an all-gather of the merged slices
                       leaves the full result on every rank: per rank (2 - 2/N) * n*k*16 B in,
                       and 1/N of the merge work;
  "allgather"          every rank receives every shard's full list and merges all queries:
                       (N - 1) * n*k*16 B in (8 GPUs, 100k queries: 112 MB against 28 MB).

torch.distributed is plumbing only: these collectives and the one-time exchange of collection
statistics at load.
"""
import os
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from .capi import HIT_DTYPE, check, lib


def combine_partition_stats(n_docs, avg_lens):
    """Collection-wide N, per-rank doc_base and average length from per-partition values. The
    same arithmetic, in the same order, on every rank, so all shards score identically."""
    n_docs = [int(x) for x in n_docs]
    total = sum(n_docs)
    bases = [sum(n_docs[:r]) for r in range(len(n_docs))]
    acc = 0.0
    for n, a in zip(n_docs, avg_lens):
        acc += float(a) * n
    return total, bases, acc / total


def merge_topk_host(hits, n_hits, k):
    """numpy statement of the cross-shard merge (checker for tests; the product path is the CUDA
    kernel): hits[shard, query, k], n_hits[shard, query] -> merged (hits[query, k], n[query]),
    ordered (score desc, doc id asc)."""
    s, n, _ = hits.shape
    out = np.zeros((n, k), HIT_DTYPE)
    out_n = np.zeros(n, np.int32)
    for q in range(n):
        rows = [hits[r, q, :n_hits[r, q]] for r in range(s)]
        allr = np.concatenate(rows) if rows else np.zeros(0, HIT_DTYPE)
        order = np.lexsort((allr["doc_id"], -allr["score"]))[:k]
        out[q, :len(order)] = allr[order]
        out_n[q] = len(order)
    return out, out_n


def slice_bounds(n, world):
    """Query slices of the scatter exchange: rank r merges queries [lo[r], lo[r + 1]); all slices
    have ceil(n / world) queries except the last non-empty one (later ranks may be empty)."""
    s = (n + world - 1) // world if n else 0
    return s, [min(n, r * s) for r in range(world + 1)]


def exchange_slices(hits_u8, n_hits_i32, n, k, rank, world, recv_hits, recv_n):
    """The all-to-all of the scatter exchange on flat tensors (uint8 hits of n*k*16 bytes, int32
    counts): afterwards recv_hits holds, shard after shard, the lists of this rank's query slice
    — the layout wsr_merge_topk_device / merge_topk_host expect. Returns the slice length."""
    _, lo = slice_bounds(n, world)
    sizes = [lo[r + 1] - lo[r] for r in range(world)]
    mine = sizes[rank]
    dist.all_to_all_single(recv_hits[:world * mine * k * 16], hits_u8[:n * k * 16],
                           output_split_sizes=[mine * k * 16] * world,
                           input_split_sizes=[x * k * 16 for x in sizes])
    dist.all_to_all_single(recv_n[:world * mine], n_hits_i32[:n],
                           output_split_sizes=[mine] * world, input_split_sizes=sizes)
    return mine


class _DevPtr:
    """Exposes a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3}


def device_bytes(ptr, nbytes, device):
    return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)


class ShardedSearch:
    """One rank of a document-partitioned deployment."""

    def __init__(self, engine, rank, world, device=None, term_keys="synthetic_rank", exchange=None):
        self.engine, self.rank, self.world = engine, rank, world
        self.exchange = exchange or os.environ.get("WSR_EXCHANGE", "scatter")
        if self.exchange not in ("scatter", "allgather"):
            raise ValueError("exchange must be 'scatter' or 'allgather'")
        self.device = device if device is not None else torch.device("cuda", engine.device)
        self.term_keys = term_keys
        self._bufs = {}
        self.exchange_stats()

    # ---- load-time exchange of collection statistics ---------------------------------------
    def exchange_stats(self):
        info = self.engine.info()
        dev = self.device if dist.get_backend() == "nccl" else torch.device("cpu")
        mine = torch.tensor([float(info.n_docs), float(info.avg_doc_len)], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allv, mine)
        n_docs = [int(v[0].item()) for v in allv]
        avgs = [float(v[1].item()) for v in allv]
        total, bases, avg = combine_partition_stats(n_docs, avgs)
        df_local, ranks = self.engine.local_stats(want_ranks=(self.term_keys == "synthetic_rank"))
        if self.term_keys == "synthetic_rank":
            vmax = torch.tensor([int(ranks.max()) + 1 if len(ranks) else 1], dtype=torch.int64, device=dev)
            dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
            dense = torch.zeros(int(vmax.item()), dtype=torch.int64, device=dev)
            dense[torch.from_numpy(ranks.astype(np.int64)).to(dev)] = torch.from_numpy(df_local.astype(np.int64)).to(dev)
            dist.all_reduce(dense)
            df_global = dense[torch.from_numpy(ranks.astype(np.int64)).to(dev)].cpu().numpy().astype(np.uint32)
        else:
            # generic vocabularies: exchange (term, df) pairs (fine for test-sized corpora)
            terms = [self.engine.term_at(i)[0] for i in range(len(df_local))]
            objs = [None] * self.world
            dist.all_gather_object(objs, dict(zip(terms, df_local.tolist())))
            tot = {}
            for o in objs:
                for t, d in o.items():
                    tot[t] = tot.get(t, 0) + d
            df_global = np.array([tot[t] for t in terms], np.uint32)
        self.doc_base, self.n_docs_global, self.avg_len_global = bases[self.rank], total, avg
        self.engine.set_global_stats(bases[self.rank], total, avg, df_global)
        return total, bases[self.rank], avg

    # ---- per-batch: exchange of the shard top-k + merge kernel -----------------------------
    def _buffers(self, n, k):
        key = (n, k)
        if key not in self._bufs:
            d = self.device
            s, _ = slice_bounds(n, self.world)
            self._bufs = {key: dict(
                g_hits=torch.empty((self.world, n * k * 16), dtype=torch.uint8, device=d),
                g_n=torch.empty((self.world, n), dtype=torch.int32, device=d),
                # scatter exchange: merged slice (padded to s queries) and the gathered slices;
                # slice r starts at query r*s, so the first n queries of full_* are the result
                m_hits=torch.zeros(max(1, s) * k * 16, dtype=torch.uint8, device=d),
                m_n=torch.zeros(max(1, s), dtype=torch.int32, device=d),
                full_hits=torch.empty(self.world * max(1, s) * k * 16, dtype=torch.uint8, device=d),
                full_n=torch.empty(self.world * max(1, s), dtype=torch.int32, device=d),
                out_hits=torch.empty(n * k * 16, dtype=torch.uint8, device=d),
                out_n=torch.empty(n, dtype=torch.int32, device=d))}
        return self._bufs[key]

    def gather_merge(self, batch):
        """Enqueued on the batch's stream, right behind its search kernels."""
        n, k = batch.n, batch.k_stride
        b = self._buffers(n, k)
        d_hits, d_n, stream_ptr = batch.device_results()
        stream = torch.cuda.ExternalStream(stream_ptr, device=self.device)
        mine_h = device_bytes(d_hits, n * k * 16, self.device)
        mine_n = device_bytes(d_n, n * 4, self.device).view(torch.int32)
        if self.exchange == "allgather":
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(b["g_hits"].view(-1), mine_h)
                dist.all_gather_into_tensor(b["g_n"].view(-1), mine_n)
            check(lib().wsr_merge_topk_device(b["g_hits"].data_ptr(), b["g_n"].data_ptr(), self.world, n, k,
                                              b["out_hits"].data_ptr(), b["out_n"].data_ptr(),
                                              C.c_void_p(stream_ptr)))
            b["res_hits"], b["res_n"] = b["out_hits"], b["out_n"]
            return b["res_hits"], b["res_n"]
        with torch.cuda.stream(stream):
            mine = exchange_slices(mine_h, mine_n, n, k, self.rank, self.world,
                                   b["g_hits"].view(-1), b["g_n"].view(-1))
        check(lib().wsr_merge_topk_device(b["g_hits"].data_ptr(), b["g_n"].data_ptr(), self.world, mine, k,
                                          b["m_hits"].data_ptr(), b["m_n"].data_ptr(),
                                          C.c_void_p(stream_ptr)))
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(b["full_hits"], b["m_hits"])
            dist.all_gather_into_tensor(b["full_n"], b["m_n"])
        b["res_hits"], b["res_n"] = b["full_hits"][:n * k * 16], b["full_n"][:n]
        return b["res_hits"], b["res_n"]

    def fetch_merged(self, batch, hits_host=None, n_host=None):
        """D2H of the merged result of the last gather_merge (synchronises the batch stream)."""
        n, k = batch.n, batch.k_stride
        b = self._buffers(n, k)
        stream = torch.cuda.ExternalStream(batch.device_results()[2], device=self.device)
        with torch.cuda.stream(stream):
            h = b["res_hits"].cpu() if hits_host is None else hits_host.copy_(b["res_hits"], non_blocking=True)
            c = b["res_n"].cpu() if n_host is None else n_host.copy_(b["res_n"], non_blocking=True)
        stream.synchronize()
        return h.numpy().view(HIT_DTYPE).reshape(n, k), c.numpy()
