"""CPU tests of the C-ABI boundary: libwsr.so loads without a GPU, exports every symbol that
include/wsr.h declares, refuses to run without a device (no CPU fallback), and the Python mirror
keeps the reference interface's names."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wiser_b200", "libwsr.so")

pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="libwsr.so not built (run __graft_entry__.build())")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wsr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wsr_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    from wiser_b200 import capi
    assert sorted(capi.EXPORTS) == syms


def test_no_cpu_fallback_without_gpu(golden_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from wiser_b200 import GpuVacuumEngine
    from wiser_b200.capi import WsrError
    with pytest.raises(WsrError, match="no CUDA device"):
        GpuVacuumEngine(os.path.join(golden_dir, "hello3")).Load()


def test_interface_mirrors_reference_names():
    import wiser_b200 as w
    q = w.SearchQuery(["a", "b"])
    assert (q.n_results, q.return_snippets, q.n_snippet_passages, q.is_phrase) == (5, False, 3, False)
    assert w.SearchQuery(["a", "b"], True).is_phrase
    for name in ("Load", "Search", "TermCount", "PostinglistSizes", "AddDocument", "LoadLocalDocuments",
                 "Serialize", "Deserialize"):
        assert hasattr(w.GpuVacuumEngine, name)
    with pytest.raises(RuntimeError, match="Wrong engine type"):
        w.CreateSearchEngine("nope")
    e = w.CreateSearchEngine("gpu:vacuum_dump:/tmp/x")
    assert isinstance(e, w.GpuVacuumEngine) and e.engine_dir_path == "/tmp/x"
    with pytest.raises(NotImplementedError):
        e.AddDocument(None)


def test_query_log_line_format():
    from wiser_b200 import parse_query_line
    q = parse_query_line(" nightt rain  nashvil \n")
    assert q.terms == ["nightt", "rain", "nashvil"] and not q.is_phrase
    q = parse_query_line('"greek armi"')
    assert q.terms == ["greek", "armi"] and q.is_phrase
    assert parse_query_line("").terms == []
