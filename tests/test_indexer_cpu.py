"""wsr_index_linedoc (native linedoc -> vacuum directory indexer) against the
golden fixtures that the REFERENCE's own dumper produced from the same linedoc input: identical
doc-length file, identical posting lists, identical query results (incl. phrase queries, i.e. the
position column) — through the CPU oracle everywhere, and through the unmodified reference
engine (oracle/_ref/ref_tool) where that binary exists."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_py import OracleIndex, parse_query_line, read_ref_results

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "wiser_b200", "wsr_index_linedoc")
REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_tool")


def _linedoc(name, tmp_path):
    p = str(tmp_path / (name + ".linedoc"))
    if name == "hello3":
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_golden
        make_golden.hello3_linedoc(p)
    else:   # zipf2k: same generator call as tests/golden/make_golden.py
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_linedoc.py"), "--docs", "2000",
                               "--vocab", "3000", "--seed", "7", "--out", p], stdout=subprocess.DEVNULL)
    return p


@pytest.mark.parametrize("name", ["hello3", "zipf2k"])
@pytest.mark.parametrize("threads", [1, 5])
def test_indexer_equals_reference_dumper(golden_dir, tmp_path, name, threads):
    if not os.path.exists(TOOL):
        pytest.skip("wsr_index_linedoc not built")
    gold = os.path.join(golden_dir, name)
    out = str(tmp_path / "idx")
    subprocess.check_call([TOOL, "--linedoc", _linedoc(name, tmp_path), "--out", out, "--threads", str(threads)],
                          stdout=subprocess.DEVNULL)
    # (1) doc lengths + running average: byte for byte
    assert open(os.path.join(out, "my.doc_length"), "rb").read() == \
        open(os.path.join(gold, "my.doc_length"), "rb").read()
    # (2) every posting list
    ix = OracleIndex(out)
    z = np.load(os.path.join(gold, "lists.npz"))
    terms, offs = z["terms"], z["offsets"]
    assert ix.term_count == len(terms)
    for i, t in enumerate(terms):
        docs, tfs = ix.decode_list(str(t))
        assert np.array_equal(docs, z["docs"][offs[i]:offs[i + 1]]), t
        assert np.array_equal(tfs, z["tfs"][offs[i]:offs[i + 1]]), t
    # (3) query results, phrase queries included (they read the position column we wrote)
    ref = read_ref_results(os.path.join(gold, "ref_top10.txt.gz"))
    qs = [parse_query_line(l) for l in open(os.path.join(gold, "queries.txt"))]
    n_phrase = 0
    for (q, is_phrase), (docs, scores, dfs) in zip(qs, ref):
        od, os_, odfs = ix.search(q, 10, is_phrase=is_phrase)
        assert np.array_equal(od, docs), q
        assert np.array_equal(os_.view(np.uint64), scores.view(np.uint64)), q
        assert odfs == dfs, q
        n_phrase += bool(is_phrase and len(q) > 1)
    assert name != "zipf2k" or n_phrase > 400
    # (4) the unmodified reference engine reads our directory and answers identically
    if os.path.exists(REF_TOOL) and threads == 1:
        got = str(tmp_path / "replay.txt")
        subprocess.check_call([REF_TOOL, "replay", out, os.path.join(gold, "queries.txt"), "10", got],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert open(got, "rb").read() == gzip.open(os.path.join(gold, "ref_top10.txt.gz"), "rb").read()


def test_indexer_rejects_malformed_rows(tmp_path):
    if not os.path.exists(TOOL):
        pytest.skip("wsr_index_linedoc not built")
    p = str(tmp_path / "bad.linedoc")
    with open(p, "w") as f:
        f.write("FIELDS_HEADER_INDICATOR###\tdoctitle\tbody\ttokenized\toffsets\tpositions\n")
        f.write("d0\ta b\ta b\t0,1;.2,3;.\t0;.\n")          # positions of 'b' missing
    r = subprocess.run([TOOL, "--linedoc", p, "--out", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode != 0 and "document 0" in r.stderr
    r = subprocess.run([TOOL, "--linedoc", str(tmp_path / "nosuch"), "--out", str(tmp_path / "o")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "File may not exist" in r.stderr
