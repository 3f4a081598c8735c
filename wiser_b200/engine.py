"""Host-side mirror of the reference's engine seam, over the C ABI (libwsr.so).

Mirrors, name for name, the reference interface this path replaces
(src/qq_mem/src/engine_services.h:14-27, types.h:205-345, engine_factory.h:33-50):
  SearchQuery, SearchResultEntry, SearchResult, SearchEngineServiceNew.{Load, Search,
  TermCount, PostinglistSizes}, CreateSearchEngine.
The C++ twin (the one a reference maintainer links) is wiser_b200/csrc/gpu_vacuum_engine.h.
Python here is test / bench plumbing; all query work happens in the CUDA library.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np

from . import capi
from .capi import HIT_DTYPE, QUERY_DTYPE, WSR_MAX_TERMS, WSR_TERM_ABSENT, check, lib


@dataclass
class SearchQuery:            # types.h:205-218 (same field names and defaults)
    terms: List[str] = field(default_factory=list)
    is_phrase: bool = False
    n_results: int = 5
    return_snippets: bool = False
    n_snippet_passages: int = 3


@dataclass
class SearchResultEntry:      # types.h:259-263
    doc_id: int = 0
    doc_score: float = 0.0
    snippet: str = ""


@dataclass
class SearchResult:           # types.h:301-303
    entries: List[SearchResultEntry] = field(default_factory=list)
    doc_freqs: List[int] = field(default_factory=list)

    def Size(self):
        return len(self.entries)

    def __getitem__(self, i):
        return self.entries[i]


def parse_query_line(line: str) -> SearchQuery:
    """One line of a query log (query_pool.h:251-311): trimmed; wrapped in double quotes means
    phrase; terms split on single spaces, empty pieces dropped (utils::explode)."""
    line = line.strip()
    is_phrase = len(line) >= 1 and line[0] == '"' and line[-1] == '"'
    if is_phrase:
        line = line[1:-1]
    return SearchQuery(terms=[t for t in line.split(" ") if t], is_phrase=is_phrase)


def load_query_log(path: str, n_results: int = 5) -> List[SearchQuery]:
    """QueryProducerByLog's loader (query_pool.h:314-335) with the engine_bench defaults
    n_results=5, return_snippets=false (types.h:215-218)."""
    out = []
    with open(path) as f:
        for line in f:
            q = parse_query_line(line.rstrip("\n"))
            q.n_results = n_results
            out.append(q)
    return out


class GpuVacuumEngine:
    """SearchEngineServiceNew over an HBM-resident vacuum index (VacuumEngine twin,
    vacuum_engine.h:120-300). URL scheme: gpu:vacuum_dump:<dir>."""

    def __init__(self, engine_dir_path: str, bloom_enable_factor: int = 1, device: int = 0,
                 shard: int = 0, n_shards: int = 1, loader_threads: int = 0, positions: bool = True):
        """positions: also load the position column (needed by phrase queries; 4 B/token)."""
        self.engine_dir_path = engine_dir_path
        self.bloom_enable_factor = bloom_enable_factor
        self.device, self.shard, self.n_shards = device, shard, n_shards
        self.loader_threads = loader_threads
        self.positions = positions
        self._h = None

    # ---- SearchEngineServiceNew ----------------------------------------------------------
    def Load(self):
        if self._h:
            raise RuntimeError("Engine is already loaded.")       # vacuum_engine.h:145
        err = C.create_string_buffer(512)
        self._h = lib().wsr_index_open_ex(self.engine_dir_path.encode(), self.device, self.shard,
                                          self.n_shards, self.loader_threads,
                                          capi.WSR_OPEN_POSITIONS if self.positions else 0, err, 512)
        if not self._h:
            raise capi.WsrError("wsr_index_open: " + err.value.decode())
        return self

    def TermCount(self) -> int:
        return int(self.info().n_terms)

    def PostinglistSizes(self, terms: Sequence[str]) -> Dict[str, int]:
        out = {}
        for t in terms:
            tid, df = self.term_lookup(t)
            if tid != WSR_TERM_ABSENT:
                out[t] = df
        return out

    def Search(self, query: SearchQuery) -> SearchResult:
        enc = [t.encode() for t in query.terms]
        n = len(enc)
        k = int(query.n_results)
        arr = (C.c_char_p * max(n, 1))(*enc)
        lens = (C.c_size_t * max(n, 1))(*[len(t) for t in enc])
        hits = np.zeros(max(k, 1), HIT_DTYPE)
        dfs = np.zeros(WSR_MAX_TERMS, np.uint32)
        nh, ndf = C.c_int(0), C.c_int(0)
        check(lib().wsr_search(self._h, arr, lens, n, k,
                               capi.WSR_QUERY_PHRASE if query.is_phrase else 0, hits.ctypes.data,
                               C.byref(nh), dfs.ctypes.data, C.byref(ndf)))
        res = SearchResult()
        for i in range(nh.value):
            res.entries.append(SearchResultEntry(int(hits["doc_id"][i]), float(hits["score"][i])))
        res.doc_freqs = [int(x) for x in dfs[:ndf.value]]
        return res

    def AddDocument(self, doc_info):
        raise NotImplementedError("Not implemented in VacuumEngine.")   # vacuum_engine.h:260-276

    def LoadLocalDocuments(self, line_doc_path, n_rows, loader):
        raise NotImplementedError("Not implemented in VacuumEngine.")

    def Serialize(self, dir_path):
        raise NotImplementedError("Not implemented in VacuumEngine.")

    def Deserialize(self, dir_path):
        raise NotImplementedError("Not implemented in VacuumEngine.")

    # ---- batch interface (what the replay driver uses) -----------------------------------
    def term_lookup(self, term: str):
        t = term.encode()
        tid, df = C.c_uint32(0), C.c_uint32(0)
        rc = lib().wsr_term_lookup(self._h, t, len(t), C.byref(tid), C.byref(df))
        if rc < 0:
            check(rc)
        return (WSR_TERM_ABSENT, 0) if rc == 1 else (tid.value, df.value)

    def make_queries(self, queries: Sequence[SearchQuery], out=None) -> np.ndarray:
        """Term lookup for a whole log -> wsr_query records."""
        arr = out if out is not None else np.zeros(len(queries), QUERY_DTYPE)
        cache = {}
        for i, q in enumerate(queries):
            if len(q.terms) > WSR_MAX_TERMS:
                raise capi.WsrError("more than WSR_MAX_TERMS terms")
            ids = arr["term_ids"][i]
            for j, t in enumerate(q.terms):
                tid = cache.get(t)
                if tid is None:
                    tid = self.term_lookup(t)[0]
                    cache[t] = tid
                ids[j] = tid
            arr["n_terms"][i] = len(q.terms)
            arr["k"][i] = q.n_results
            arr["flags"][i] = capi.WSR_QUERY_PHRASE if q.is_phrase else 0
        return arr

    def parse_query_log(self, text: bytes, k: int) -> np.ndarray:
        """wsr_parse_query_log: raw log text -> wsr_query records (term lookup in C)."""
        cap = text.count(b"\n") + 2
        arr = np.zeros(cap, QUERY_DTYPE)
        n = C.c_int(0)
        check(lib().wsr_parse_query_log(self._h, text, len(text), k, arr.ctypes.data, cap,
                                        C.byref(n)))
        return arr[:n.value]

    def search_batch(self, qarr: np.ndarray, k_stride: int, hits=None, n_hits=None,
                     want_doc_freqs=False):
        """wsr_search_batch over host buffers -> (hits[n,k_stride], n_hits[n], dfs, n_dfs)."""
        n = len(qarr)
        if hits is None:
            hits = np.zeros((n, k_stride), HIT_DTYPE)
        if n_hits is None:
            n_hits = np.zeros(n, np.int32)
        dfs = np.zeros((n, WSR_MAX_TERMS), np.uint32) if want_doc_freqs else None
        ndfs = np.zeros(n, np.int32) if want_doc_freqs else None
        check(lib().wsr_search_batch(self._h, qarr.ctypes.data, n, k_stride, hits.ctypes.data,
                                     n_hits.ctypes.data,
                                     dfs.ctypes.data if want_doc_freqs else None,
                                     ndfs.ctypes.data if want_doc_freqs else None))
        return hits, n_hits, dfs, ndfs

    def search_log(self, text: bytes, k: int, hits=None, n_hits=None, doc_freqs=None, n_doc_freqs=None):
        """wsr_search_log: whole query-log text -> (hits[n,k], n_hits[n]); parse, GPU and copies
        pipelined inside the library."""
        if isinstance(text, np.ndarray):       # e.g. a pinned uint8 buffer holding the log
            ptr, length = C.c_void_p(text.ctypes.data), int(text.size)
            cap = len(n_hits) if n_hits is not None else int(np.count_nonzero(text == 10)) + 2
        else:
            ptr, length = C.cast(C.c_char_p(text), C.c_void_p), len(text)
            cap = text.count(b"\n") + 2
        if hits is None:
            hits = np.zeros((cap, k), HIT_DTYPE)
        if n_hits is None:
            n_hits = np.zeros(cap, np.int32)
        n = C.c_int(0)
        if doc_freqs is not None:
            check(lib().wsr_search_log_ex(self._h, ptr, length, k, hits.ctypes.data, n_hits.ctypes.data,
                                          doc_freqs.ctypes.data, n_doc_freqs.ctypes.data,
                                          min(cap, len(n_hits)), C.byref(n)))
        else:
            check(lib().wsr_search_log(self._h, ptr, length, k, hits.ctypes.data, n_hits.ctypes.data,
                                       min(cap, len(n_hits)), C.byref(n)))
        return hits[:n.value], n_hits[:n.value]

    def SearchBatch(self, queries: Sequence[SearchQuery]) -> List[SearchResult]:
        if not queries:
            return []
        qarr = self.make_queries(queries)
        k_stride = max(1, max(q.n_results for q in queries))
        hits, n_hits, dfs, ndfs = self.search_batch(qarr, k_stride, want_doc_freqs=True)
        out = []
        for i in range(len(queries)):
            r = SearchResult()
            for j in range(int(n_hits[i])):
                r.entries.append(SearchResultEntry(int(hits["doc_id"][i, j]),
                                                   float(hits["score"][i, j])))
            r.doc_freqs = [int(x) for x in dfs[i, :ndfs[i]]]
            out.append(r)
        return out

    def local_stats(self, want_ranks=False):
        """(df_local[n_terms], ranks[n_terms] | None) for the document-partition stats exchange."""
        n = self.TermCount()
        df = np.zeros(max(n, 1), np.uint32)
        ranks = np.zeros(max(n, 1), np.uint32) if want_ranks else None
        check(lib().wsr_index_local_stats(self._h, df.ctypes.data,
                                          ranks.ctypes.data if want_ranks else None))
        return df[:n], (ranks[:n] if want_ranks else None)

    def set_global_stats(self, doc_base: int, n_docs_global: int, avg_len_global: float,
                         df_global: np.ndarray):
        df_global = np.ascontiguousarray(df_global, np.uint32)
        assert len(df_global) == self.TermCount()
        check(lib().wsr_index_set_global_stats(self._h, int(doc_base), int(n_docs_global),
                                               float(avg_len_global), df_global.ctypes.data))

    def decode_list(self, term: str):
        tid, _ = self.term_lookup(term)
        if tid == WSR_TERM_ABSENT:
            return None
        n = C.c_size_t(0)
        check(lib().wsr_decode_list(self._h, tid, None, None, 0, C.byref(n)))
        docs = np.zeros(max(n.value, 1), np.uint32)
        tfs = np.zeros(max(n.value, 1), np.uint32)
        check(lib().wsr_decode_list(self._h, tid, docs.ctypes.data, tfs.ctypes.data, n.value,
                                    C.byref(n)))
        return docs[:n.value], tfs[:n.value]

    def decode_all(self):
        s, ms = C.c_uint64(0), C.c_float(0)
        check(lib().wsr_decode_all(self._h, C.byref(s), C.byref(ms)))
        return s.value, ms.value

    def info(self) -> capi.IndexInfo:
        inf = capi.IndexInfo()
        check(lib().wsr_index_get_info(self._h, C.byref(inf)))
        return inf

    def term_at(self, term_id: int):
        buf = C.create_string_buffer(4096)
        df = C.c_uint32(0)
        n = lib().wsr_term_at(self._h, term_id, buf, 4096, C.byref(df))
        if n < 0:
            check(n)
        return buf.raw[:n].decode(), df.value

    def close(self):
        if self._h and not getattr(self, "_borrowed", False):   # a group's partitions belong to the group
            lib().wsr_index_close(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """Device-resident query batch (wsr_batch_*): upload once, run many times."""

    def __init__(self, engine: GpuVacuumEngine, qarr: np.ndarray, k_stride: int):
        self.engine = engine
        self.n = len(qarr)
        self.k_stride = k_stride
        self._b = lib().wsr_batch_create(engine._h, qarr.ctypes.data, self.n, k_stride)
        if not self._b:
            raise capi.WsrError("wsr_batch_create: " + lib().wsr_last_error().decode())

    def reset(self, qarr: np.ndarray, k_stride: int):
        self.n, self.k_stride = len(qarr), k_stride
        check(lib().wsr_batch_reset(self._b, qarr.ctypes.data, self.n, k_stride))

    def reset_log(self, text, k: int) -> int:
        """wsr_batch_reset_log: re-plan this batch from query-log text (bytes or a pinned uint8
        array); parse, term lookup and planning run on the GPU. Returns the number of queries."""
        if isinstance(text, np.ndarray):
            ptr, length = C.c_void_p(text.ctypes.data), int(text.size)
        else:
            ptr, length = C.cast(C.c_char_p(text), C.c_void_p), len(text)
        n = C.c_int(0)
        check(lib().wsr_batch_reset_log(self._b, ptr, length, k, C.byref(n)))
        self.n, self.k_stride = n.value, k
        return n.value

    def run(self):
        check(lib().wsr_batch_run(self._b))

    def sync(self):
        check(lib().wsr_batch_sync(self._b))

    def time(self, iters: int) -> float:
        ms = C.c_float(0)
        check(lib().wsr_batch_time(self._b, iters, C.byref(ms)))
        return ms.value

    def fetch(self, hits=None, n_hits=None):
        if hits is None:
            hits = np.zeros((self.n, self.k_stride), HIT_DTYPE)
        if n_hits is None:
            n_hits = np.zeros(self.n, np.int32)
        check(lib().wsr_batch_fetch(self._b, hits.ctypes.data, n_hits.ctypes.data))
        return hits, n_hits

    def profile(self):
        ms = (C.c_float * 6)()
        check(lib().wsr_batch_profile(self._b, C.byref(ms)))
        return [float(x) for x in ms]

    def count_work(self):
        """One pass through the counting kernel instantiations; stats() then holds its counters."""
        check(lib().wsr_batch_count_work(self._b))

    def stats(self) -> capi.BatchStats:
        s = capi.BatchStats()
        check(lib().wsr_batch_get_stats(self._b, C.byref(s)))
        return s

    def device_results(self):
        h, n, st = C.c_void_p(0), C.c_void_p(0), C.c_void_p(0)
        check(lib().wsr_batch_device_results(self._b, C.byref(h), C.byref(n), C.byref(st)))
        return h.value, n.value, st.value

    def close(self):
        if self._b:
            lib().wsr_batch_destroy(self._b)
            self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def CreateSearchEngine(engine_type: str, bloom_enable_factor: int = 1, **kw) -> GpuVacuumEngine:
    """engine_factory.h:33-50 with the new URL scheme gpu:vacuum_dump:<dir>."""
    parts = engine_type.split(":")
    if len(parts) == 3 and parts[0] == "gpu" and parts[1] == "vacuum_dump":
        return GpuVacuumEngine(parts[2], bloom_enable_factor, **kw)
    raise RuntimeError("Wrong engine type: " + engine_type)
