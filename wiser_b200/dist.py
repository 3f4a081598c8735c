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

from . import capi
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


def _new_comm_id(rank):
    """NCCL unique id for the C-ABI exchange: created by rank 0 in the library, handed to the
    other ranks through torch.distributed (plumbing)."""
    buf = C.create_string_buffer(capi.WSR_COMM_ID_BYTES)
    if rank == 0:
        check(lib().wsr_comm_unique_id(buf))
    objs = [buf.raw if rank == 0 else None]
    dist.broadcast_object_list(objs, src=0)
    return objs[0]


def _exchange_df(df_locals, ranks_list, dev):
    """Collection-wide df of every local term. Synthetic vocabularies (terms named t<rank>): a
    dense all-reduce over term ranks; general ones: an exchange of (term, df) tables."""
    vmax = torch.tensor([max([int(r.max()) + 1 if len(r) else 1 for r in ranks_list])], dtype=torch.int64, device=dev)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
    dense = torch.zeros(int(vmax.item()), dtype=torch.int64, device=dev)
    for df, r in zip(df_locals, ranks_list):
        dense.index_add_(0, torch.from_numpy(r.astype(np.int64)).to(dev), torch.from_numpy(df.astype(np.int64)).to(dev))
    dist.all_reduce(dense)
    return [dense[torch.from_numpy(r.astype(np.int64)).to(dev)].cpu().numpy().astype(np.uint32) for r in ranks_list]


class ShardedSearch:
    """One rank of a document-partitioned deployment: one engine (one partition directory) per
    process. The per-batch exchange runs in the library (wsr_batch_exchange: NCCL called from C on
    the batch's stream); torch.distributed only carries the communicator id and the one-time
    exchange of collection statistics."""

    def __init__(self, engine, rank, world, device=None, term_keys="synthetic_rank", exchange=None):
        if getattr(engine, "n_shards", 1) != 1:
            # a doc-range shard of ONE directory already scores with the collection's statistics
            # and emits global doc ids; feeding it through the partition exchange would count N
            # world times and offset the ids twice
            raise ValueError("ShardedSearch expects every rank to open its OWN partition directory "
                             "(n_shards == 1); doc-range shards of one directory need no statistics exchange")
        self.engine, self.rank, self.world = engine, rank, world
        self.exchange = exchange or os.environ.get("WSR_EXCHANGE", "scatter")
        if self.exchange not in ("scatter", "allgather"):
            raise ValueError("exchange must be 'scatter' or 'allgather'")
        self.device = device if device is not None else torch.device("cuda", engine.device)
        self.term_keys = term_keys
        self._comm = None
        self.exchange_stats()
        if dist.get_backend() == "nccl":
            cid = _new_comm_id(rank)
            self._comm = lib().wsr_comm_init_rank(cid, rank, world, engine.device)
            if not self._comm:
                raise capi.WsrError("wsr_comm_init_rank: " + lib().wsr_last_error().decode())

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
            df_global = _exchange_df([df_local], [ranks], dev)[0]
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

    # ---- per-batch: exchange of the shard top-k + merge kernel, in the library -------------
    def gather_merge(self, batch):
        """Enqueued on the batch's stream, right behind its search kernels."""
        check(lib().wsr_batch_exchange(batch._b, self._comm, 0 if self.exchange == "scatter" else 1))

    def fetch_merged(self, batch, hits_host=None, n_host=None):
        """D2H of the merged result of the last gather_merge (synchronises the batch stream).
        hits_host / n_host: optional pinned numpy arrays (capi.PinnedArray.array)."""
        n, k = batch.n, batch.k_stride
        h = hits_host if hits_host is not None else np.zeros((n, k), HIT_DTYPE)
        c = n_host if n_host is not None else np.zeros(n, np.int32)
        check(lib().wsr_batch_fetch_exchanged(batch._b, self._comm, h.ctypes.data, c.ctypes.data))
        return h.reshape(-1)[:n * k].reshape(n, k), c[:n]

    def close(self):
        if self._comm:
            lib().wsr_comm_destroy(self._comm)
            self._comm = None


class ShardGroup:
    """Binding of wsr_group (include/wsr.h): the partitions this process holds — one or more
    partition directories on one or more of its GPUs — as one engine. Single process: the library
    does everything (statistics exchange between partitions, NCCL between devices). One process
    per GPU (rank / world given): the communicator id and the collection statistics travel through
    torch.distributed once at load; every batch is search kernels + on-device merge of the local
    partitions + NCCL exchange called from C."""

    def __init__(self, dirs, devices, rank=None, world=None, positions=False, loader_threads=0):
        self.dirs, self.devices = list(dirs), list(devices)
        self.rank, self.world = rank, world
        multi = world is not None and world > 1
        arr = (C.c_char_p * len(self.dirs))(*[d.encode() for d in self.dirs])
        devs = (C.c_int * len(self.devices))(*self.devices)
        gd = None
        if multi:
            gd = capi.GroupDist()
            gd.rank, gd.world = rank, world
            cid = _new_comm_id(rank)          # 128 raw bytes (may contain NULs: no string assignment)
            C.memmove(C.addressof(gd) + capi.GroupDist.comm_id.offset, cid, capi.WSR_COMM_ID_BYTES)
        err = C.create_string_buffer(512)
        self._g = lib().wsr_group_open(arr, len(self.dirs), devs, len(self.devices), loader_threads,
                                       capi.WSR_OPEN_POSITIONS if positions else 0,
                                       C.byref(gd) if gd is not None else None, err, 512)
        if not self._g:
            raise capi.WsrError("wsr_group_open: " + err.value.decode())
        self.n = self.k = 0
        if multi:
            self._exchange_stats()

    def part(self, i):
        """Partition i as an engine object over the group's own index (borrowed: closing it is a no-op)."""
        from .engine import GpuVacuumEngine
        e = GpuVacuumEngine(self.dirs[i], device=self.devices[i // max(1, len(self.dirs) // len(self.devices))])
        e._h = lib().wsr_group_part(self._g, i)
        e._borrowed = True
        return e

    def _exchange_stats(self):
        dev = torch.device("cuda", self.devices[0])
        parts = [self.part(i) for i in range(len(self.dirs))]
        infos = [p.info() for p in parts]
        mine = torch.tensor([[float(i.n_docs), float(i.avg_doc_len)] for i in infos], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allv, mine)
        n_docs = [int(v[p, 0].item()) for v in allv for p in range(len(parts))]
        avgs = [float(v[p, 1].item()) for v in allv for p in range(len(parts))]
        total, bases, avg = combine_partition_stats(n_docs, avgs)
        stats = [p.local_stats(want_ranks=True) for p in parts]
        dfg = _exchange_df([s[0] for s in stats], [s[1] for s in stats], dev)
        for i, p in enumerate(parts):
            p.set_global_stats(bases[self.rank * len(parts) + i], total, avg, dfg[i])
        self.n_docs_global, self.avg_len_global = total, avg

    def load_log(self, text, k):
        ptr, length = _text_ptr(text)
        n = C.c_int(0)
        check(lib().wsr_group_load_log(self._g, ptr, length, k, C.byref(n)))
        self.n, self.k = n.value, k
        return n.value

    def run(self, mode="scatter"):
        check(lib().wsr_group_run(self._g, 0 if mode == "scatter" else 1))

    def sync(self):
        check(lib().wsr_group_sync(self._g))

    def join(self):
        """The leading stream waits for the last (pipelined) exchange: call before an end event."""
        check(lib().wsr_group_join(self._g))

    def stream(self):
        s = C.c_void_p(0)
        check(lib().wsr_group_stream(self._g, C.byref(s)))
        return s.value

    def fetch(self, hits=None, n_hits=None):
        h = hits if hits is not None else np.zeros((self.n, self.k), HIT_DTYPE)
        c = n_hits if n_hits is not None else np.zeros(self.n, np.int32)
        check(lib().wsr_group_fetch(self._g, h.ctypes.data, c.ctypes.data))
        return h.reshape(-1)[:self.n * self.k].reshape(self.n, self.k), c[:self.n]

    def search_log(self, text, k, hits=None, n_hits=None, doc_freqs=None, n_doc_freqs=None, fetch=True):
        """wsr_group_search_log. fetch=False: a rank that does not face the client."""
        ptr, length = _text_ptr(text)
        n = C.c_int(0)
        if not fetch:
            check(lib().wsr_group_search_log(self._g, ptr, length, k, None, None, None, None, 0, C.byref(n)))
            return n.value
        cap = len(n_hits)
        check(lib().wsr_group_search_log(self._g, ptr, length, k, hits.ctypes.data, n_hits.ctypes.data,
                                         doc_freqs.ctypes.data if doc_freqs is not None else None,
                                         n_doc_freqs.ctypes.data if n_doc_freqs is not None else None,
                                         cap, C.byref(n)))
        return n.value

    def stats(self):
        a, b, c, d = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_int64(0)
        check(lib().wsr_group_stats(self._g, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"listed_postings": a.value, "n_postings": b.value, "hbm_bytes": c.value, "n_docs": d.value}

    def close(self):
        if self._g:
            lib().wsr_group_close(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _text_ptr(text):
    if isinstance(text, np.ndarray):
        return C.c_void_p(text.ctypes.data), int(text.size)
    return C.cast(C.c_char_p(text), C.c_void_p), len(text)
