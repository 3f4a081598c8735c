"""wiser_b200 — B200-native engine for WiSER/Vacuum's conjunctive BM25 top-k hot path.

The product is the C-ABI CUDA library (include/wsr.h, wiser_b200/libwsr.so) plus the C++
adapter in wiser_b200/csrc/gpu_vacuum_engine.h; this package is the Python mirror of the
reference's engine interface used by the tests and bench.py."""
from .engine import (Batch, CreateSearchEngine, GpuVacuumEngine, SearchQuery, SearchResult,
                     SearchResultEntry, load_query_log, parse_query_line)

__all__ = ["Batch", "CreateSearchEngine", "GpuVacuumEngine", "SearchQuery", "SearchResult",
           "SearchResultEntry", "load_query_log", "parse_query_line"]
