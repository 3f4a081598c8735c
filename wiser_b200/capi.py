"""ctypes binding of the C ABI in include/wsr.h (libwsr.so, built in-tree by
wiser_b200/csrc/Makefile). There is no fallback: if the CUDA library is missing or no GPU is
present, loading / opening an index raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwsr.so")

WSR_MAX_TERMS = 8
WSR_TERM_ABSENT = 0xFFFFFFFF
WSR_QUERY_PHRASE = 1
WSR_OPEN_POSITIONS = 1

QUERY_DTYPE = np.dtype([("term_ids", np.uint32, (WSR_MAX_TERMS,)), ("n_terms", np.uint32),
                        ("k", np.uint32), ("flags", np.uint32)])
HIT_DTYPE = np.dtype([("doc_id", np.int32), ("reserved", np.int32), ("score", np.float64)])
assert QUERY_DTYPE.itemsize == 44 and HIT_DTYPE.itemsize == 16


class IndexInfo(C.Structure):
    _fields_ = [("n_docs", C.c_int64), ("avg_doc_len", C.c_double), ("n_terms", C.c_int64),
                ("n_postings", C.c_int64), ("n_postings_global", C.c_int64),
                ("n_blocks", C.c_int64), ("hbm_bytes", C.c_int64), ("payload_bytes", C.c_int64),
                ("shard", C.c_int32), ("n_shards", C.c_int32), ("doc_lo", C.c_int32),
                ("doc_hi", C.c_int32), ("device", C.c_int32)]


class BatchStats(C.Structure):
    _fields_ = [("listed_postings", C.c_uint64), ("decoded_postings", C.c_uint64),
                ("touched_bytes", C.c_uint64), ("listed_bytes", C.c_uint64),
                ("matches", C.c_uint64), ("work_units", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32),
                ("probe_blocks", C.c_uint64)]


EXPORTS = [
    "wsr_last_error", "wsr_device_count", "wsr_host_alloc", "wsr_host_free", "wsr_index_open", "wsr_index_open_ex",
    "wsr_index_close", "wsr_index_get_info", "wsr_term_lookup", "wsr_term_at", "wsr_decode_list",
    "wsr_decode_all", "wsr_search", "wsr_search_batch", "wsr_batch_create", "wsr_batch_destroy",
    "wsr_batch_run", "wsr_batch_sync", "wsr_batch_fetch", "wsr_batch_device_results",
    "wsr_batch_time", "wsr_batch_get_stats", "wsr_merge_topk_device", "wsr_batch_profile", "wsr_batch_count_work", "wsr_batch_reset_log",
    "wsr_parse_query_log", "wsr_index_set_global_stats", "wsr_index_local_stats", "wsr_batch_reset", "wsr_search_log",
    "wsr_search_log_ex", "wsr_comm_unique_id", "wsr_comm_init_rank", "wsr_comm_destroy", "wsr_batch_exchange",
    "wsr_batch_exchanged_results", "wsr_batch_fetch_exchanged", "wsr_group_open", "wsr_group_close",
    "wsr_group_n_parts", "wsr_group_part", "wsr_group_search_log", "wsr_group_load_log", "wsr_group_run",
    "wsr_group_sync", "wsr_group_join", "wsr_group_stream", "wsr_group_fetch", "wsr_group_stats",
]

WSR_COMM_ID_BYTES = 128


class GroupDist(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("comm_id", C.c_char * WSR_COMM_ID_BYTES)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C wiser_b200/csrc` "
                           "(python -c 'import __graft_entry__ as g; g.build()'); there is no "
                           "CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
    L.wsr_last_error.restype = cp
    L.wsr_host_alloc.restype = vp
    L.wsr_host_alloc.argtypes = [sz]
    L.wsr_host_free.argtypes = [vp]
    L.wsr_index_open.restype = vp
    L.wsr_index_open.argtypes = [cp, C.c_int, C.c_int, C.c_int, C.c_int, cp, sz]
    L.wsr_index_open_ex.restype = vp
    L.wsr_index_open_ex.argtypes = [cp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, cp, sz]
    L.wsr_index_close.argtypes = [vp]
    L.wsr_index_get_info.argtypes = [vp, C.POINTER(IndexInfo)]
    L.wsr_term_lookup.argtypes = [vp, cp, sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.wsr_term_at.argtypes = [vp, C.c_uint32, cp, sz, C.POINTER(C.c_uint32)]
    L.wsr_decode_list.argtypes = [vp, C.c_uint32, vp, vp, sz, C.POINTER(sz)]
    L.wsr_decode_all.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_float)]
    L.wsr_search.argtypes = [vp, C.POINTER(cp), C.POINTER(sz), C.c_int, C.c_int, C.c_uint, vp,
                             C.POINTER(C.c_int), vp, C.POINTER(C.c_int)]
    L.wsr_search_batch.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]
    L.wsr_batch_create.restype = vp
    L.wsr_batch_create.argtypes = [vp, vp, C.c_int, C.c_int]
    L.wsr_batch_destroy.argtypes = [vp]
    L.wsr_batch_run.argtypes = [vp]
    L.wsr_batch_sync.argtypes = [vp]
    L.wsr_batch_fetch.argtypes = [vp, vp, vp]
    L.wsr_batch_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.wsr_batch_time.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.wsr_batch_get_stats.argtypes = [vp, C.POINTER(BatchStats)]
    L.wsr_batch_profile.argtypes = [vp, C.POINTER(C.c_float * 6)]
    L.wsr_batch_count_work.argtypes = [vp]
    L.wsr_batch_reset_log.argtypes = [vp, vp, sz, C.c_int, C.POINTER(C.c_int)]
    L.wsr_parse_query_log.argtypes = [vp, cp, sz, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
    L.wsr_index_set_global_stats.argtypes = [vp, C.c_int64, C.c_int64, C.c_double, vp]
    L.wsr_index_local_stats.argtypes = [vp, vp, vp]
    L.wsr_batch_reset.argtypes = [vp, vp, C.c_int, C.c_int]
    L.wsr_search_log.argtypes = [vp, vp, sz, C.c_int, vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.wsr_merge_topk_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.wsr_search_log_ex.argtypes = [vp, vp, sz, C.c_int, vp, vp, vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.wsr_comm_unique_id.argtypes = [vp]
    L.wsr_comm_init_rank.restype = vp
    L.wsr_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.wsr_comm_destroy.argtypes = [vp]
    L.wsr_batch_exchange.argtypes = [vp, vp, C.c_int]
    L.wsr_batch_exchanged_results.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp)]
    L.wsr_batch_fetch_exchanged.argtypes = [vp, vp, vp, vp]
    L.wsr_group_open.restype = vp
    L.wsr_group_open.argtypes = [C.POINTER(cp), C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_uint,
                                 C.POINTER(GroupDist), cp, sz]
    L.wsr_group_close.argtypes = [vp]
    L.wsr_group_n_parts.argtypes = [vp]
    L.wsr_group_part.restype = vp
    L.wsr_group_part.argtypes = [vp, C.c_int]
    L.wsr_group_search_log.argtypes = [vp, vp, sz, C.c_int, vp, vp, vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.wsr_group_load_log.argtypes = [vp, vp, sz, C.c_int, C.POINTER(C.c_int)]
    L.wsr_group_run.argtypes = [vp, C.c_int]
    L.wsr_group_sync.argtypes = [vp]
    L.wsr_group_join.argtypes = [vp]
    L.wsr_group_stream.argtypes = [vp, C.POINTER(vp)]
    L.wsr_group_fetch.argtypes = [vp, vp, vp]
    L.wsr_group_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_int64)]
    _lib = L
    return L


class WsrError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise WsrError(f"libwsr error {rc}: {lib().wsr_last_error().decode()}")


class PinnedArray:
    """numpy view over page-locked memory from wsr_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self._p = lib().wsr_host_alloc(max(n, 1))
        if not self._p:
            raise WsrError("wsr_host_alloc failed: " + lib().wsr_last_error().decode())
        buf = (C.c_char * max(n, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        if getattr(self, "_p", None):
            self.array = None
            lib().wsr_host_free(self._p)
            self._p = None
