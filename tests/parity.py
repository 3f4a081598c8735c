"""Parity checker shared by the GPU tests, smoke() and bench.py.

Bar (BASELINE.json north_star): bit-exact intersection doc-id sets, BM25 scores within 1e-5
relative (we assert BIT-EXACT fp64), identical top-k except ties. The reference's choice among
equal scores is implementation-defined (heap order, SURVEY §7 "Ties"), so for a tied score the
device's doc must merely belong to the docs having that score in the full intersection."""
import numpy as np


def check_topk(ref_docs, ref_scores, got_docs, got_scores, full_docs, full_scores, what=""):
    ref_scores = np.asarray(ref_scores, np.float64)
    got_scores = np.asarray(got_scores, np.float64)
    assert len(got_docs) == len(ref_docs), f"{what}: {len(got_docs)} hits, reference {len(ref_docs)}"
    assert np.array_equal(got_scores.view(np.uint64), ref_scores.view(np.uint64)), \
        f"{what}: score vectors differ\n got {got_scores}\n ref {ref_scores}"
    assert len(set(int(d) for d in got_docs)) == len(got_docs), f"{what}: duplicate docs"
    by_score = {}
    for d, s in zip(full_docs, full_scores):
        by_score.setdefault(float(s), set()).add(int(d))
    for i, (d, s) in enumerate(zip(got_docs, got_scores)):
        docs_at = by_score.get(float(s))
        assert docs_at is not None and int(d) in docs_at, f"{what}: doc {d} score {s} not in intersection"
        if len(docs_at) == 1:
            assert int(d) == int(ref_docs[i]), f"{what}: rank {i} doc {d} != reference {ref_docs[i]}"
    # device order is deterministic: score desc, doc id asc
    for i in range(1, len(got_docs)):
        assert got_scores[i - 1] > got_scores[i] or (got_scores[i - 1] == got_scores[i] and
                                                     got_docs[i - 1] < got_docs[i]), f"{what}: order"


def check_full(ref_docs, ref_scores, got_docs, got_scores, what=""):
    """Full intersection: same doc-id set, bit-exact score per doc."""
    ro = np.argsort(np.asarray(ref_docs), kind="stable")
    go = np.argsort(np.asarray(got_docs), kind="stable")
    assert np.array_equal(np.asarray(ref_docs)[ro], np.asarray(got_docs)[go]), f"{what}: doc sets differ"
    a = np.asarray(ref_scores, np.float64)[ro].view(np.uint64)
    b = np.asarray(got_scores, np.float64)[go].view(np.uint64)
    assert np.array_equal(a, b), f"{what}: scores differ"
