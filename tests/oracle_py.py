"""ctypes binding of oracle/libwsr_oracle.so — TEST INFRASTRUCTURE (checker only).

Nothing under wiser_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libwsr_oracle.so")
REF_TOOL = os.path.join(ORACLE_DIR, "_ref", "ref_tool")


def build_oracle():
    src = os.path.join(ORACLE_DIR, "wsr_oracle.cc")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.wsr_oracle_open.restype = C.c_void_p
        L.wsr_oracle_open.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.wsr_oracle_close.argtypes = [C.c_void_p]
        L.wsr_oracle_num_docs.argtypes = [C.c_void_p]
        L.wsr_oracle_avg_doc_len.argtypes = [C.c_void_p]
        L.wsr_oracle_avg_doc_len.restype = C.c_double
        L.wsr_oracle_term_count.argtypes = [C.c_void_p]
        L.wsr_oracle_set_global_stats.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.wsr_oracle_set_global_df.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int32]
        L.wsr_oracle_term_at.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.wsr_oracle_term_df.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.wsr_oracle_term_df.restype = C.c_int64
        L.wsr_oracle_norm_byte.argtypes = [C.c_void_p, C.c_int]
        L.wsr_oracle_decode_list.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p,
                                             C.c_void_p, C.c_size_t]
        L.wsr_oracle_decode_list.restype = C.c_int64
        L.wsr_oracle_search.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t),
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
        L.wsr_oracle_search_ex.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t),
                                           C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                           C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
        L.wsr_oracle_time_batch.argtypes = [C.c_void_p, C.POINTER(C.c_char_p),
                                            C.POINTER(C.c_size_t), C.c_void_p, C.c_int64, C.c_int,
                                            C.c_int, C.POINTER(C.c_uint64)]
        L.wsr_oracle_time_batch.restype = C.c_double
        _lib = L
    return _lib


class OracleIndex:
    """CPU oracle over a vacuum index directory."""

    def __init__(self, vacuum_dir):
        err = C.create_string_buffer(256)
        self._h = lib().wsr_oracle_open(vacuum_dir.encode(), err, 256)
        if not self._h:
            raise RuntimeError("oracle open failed: " + err.value.decode())

    def close(self):
        if self._h:
            lib().wsr_oracle_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @property
    def num_docs(self):
        return lib().wsr_oracle_num_docs(self._h)

    @property
    def avg_doc_len(self):
        return lib().wsr_oracle_avg_doc_len(self._h)

    @property
    def term_count(self):
        return lib().wsr_oracle_term_count(self._h)

    def terms(self):
        buf = C.create_string_buffer(4096)
        out = []
        for i in range(self.term_count):
            n = lib().wsr_oracle_term_at(self._h, i, buf, 4096)
            out.append(buf.raw[:n].decode())
        return out

    def set_global_stats(self, n_docs_global, avg_len_global):
        """Partition mode: score with the collection's N and average length."""
        if lib().wsr_oracle_set_global_stats(self._h, int(n_docs_global), float(avg_len_global)) != 0:
            raise ValueError("bad global statistics")

    def set_global_df(self, term, df_global):
        t = term.encode()
        return lib().wsr_oracle_set_global_df(self._h, t, len(t), int(df_global)) == 0

    def df(self, term):
        t = term.encode()
        return lib().wsr_oracle_term_df(self._h, t, len(t))

    def norm_byte(self, doc):
        return lib().wsr_oracle_norm_byte(self._h, doc)

    def decode_list(self, term):
        t = term.encode()
        df = self.df(term)
        if df < 0:
            return None
        docs = np.empty(df, np.uint32)
        tfs = np.empty(df, np.uint32)
        lib().wsr_oracle_decode_list(self._h, t, len(t), docs.ctypes.data, tfs.ctypes.data, df)
        return docs, tfs

    def search(self, terms, k=10, is_phrase=False):
        """-> (docs[int32], scores[float64], doc_freqs[list]) exactly as VacuumEngine::Search."""
        enc = [t.encode() for t in terms]
        n = len(enc)
        arr = (C.c_char_p * max(n, 1))(*enc)
        lens = (C.c_size_t * max(n, 1))(*[len(t) for t in enc])
        cap = max(k, 0)
        if cap > (1 << 24):
            cap = max(1, min([self.df(t) for t in terms if self.df(t) >= 0] or [1]))
        docs = np.empty(max(cap, 1), np.int32)
        scores = np.empty(max(cap, 1), np.float64)
        dfs = np.empty(max(n, 1), np.int32)
        nh, ndf = C.c_int(0), C.c_int(0)
        rc = lib().wsr_oracle_search_ex(self._h, arr, lens, n, k, int(is_phrase), docs.ctypes.data,
                                        scores.ctypes.data, cap, C.byref(nh), dfs.ctypes.data,
                                        C.byref(ndf))
        if rc != 0:
            raise RuntimeError(f"oracle search rc={rc}")
        return docs[:nh.value].copy(), scores[:nh.value].copy(), dfs[:ndf.value].tolist()


def partitioned_search(oracles, doc_bases, terms, k, is_phrase=False):
    """One query against document partitions (OracleIndex objects in partition mode, local doc
    ids): every partition's top-k and FULL match list with global doc ids, concatenated. The
    caller picks the global top-k with the tie-aware checker (tests/parity.py)."""
    docs, scores, dfs = [], [], None
    for ora, base in zip(oracles, doc_bases):
        d, s, f = ora.search(terms, k, is_phrase=is_phrase)
        docs.append(d.astype(np.int64) + base)
        scores.append(s)
        if f:
            dfs = f
    docs = np.concatenate(docs) if docs else np.zeros(0, np.int64)
    scores = np.concatenate(scores) if scores else np.zeros(0)
    order = np.lexsort((docs, -scores))
    return docs[order].astype(np.int32), scores[order], dfs or []


def parse_query_line(line):
    """query_pool.h:251-311: trim; a line wrapped in double quotes is a phrase; split on ' '
    (utils::explode drops empty pieces)."""
    line = line.strip()
    is_phrase = len(line) >= 1 and line.startswith('"') and line.endswith('"')
    if is_phrase:
        line = line[1:-1]
    return [t for t in line.split(" ") if t], is_phrase


def read_ref_results(path):
    """Parse ref_tool replay output -> list of (docs, scores, dfs)."""
    import gzip
    out = []
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        for line in f:
            it = line.split()
            ne, nd = int(it[0]), int(it[1])
            docs = np.array([int(it[2 + 2 * i]) for i in range(ne)], np.int32)
            scores = np.array([float.fromhex(it[3 + 2 * i]) for i in range(ne)], np.float64)
            dfs = [int(x) for x in it[2 + 2 * ne: 2 + 2 * ne + nd]]
            out.append((docs, scores, dfs))
    return out
