"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against
the CPU oracle and the committed reference outputs on the golden fixtures."""
import os

import numpy as np
import pytest

from oracle_py import OracleIndex, parse_query_line, read_ref_results
from parity import check_full, check_topk

pytestmark = pytest.mark.gpu

FIXTURES = ["hello3", "abc3", "wiki4", "zipf2k"]


@pytest.fixture(scope="module")
def engines(golden_dir):
    from wiser_b200 import GpuVacuumEngine
    out = {}
    for name in FIXTURES:
        d = os.path.join(golden_dir, name)
        out[name] = (GpuVacuumEngine(d).Load(), OracleIndex(d), d)
    yield out
    for e, _, _ in out.values():
        e.close()


@pytest.mark.parametrize("name", FIXTURES)
def test_decode_every_list(engines, name):
    """K1: every posting list decodes to the (doc id, tf) sequence of the reference iterators."""
    eng, _, d = engines[name]
    z = np.load(os.path.join(d, "lists.npz"))
    terms, offs = z["terms"], z["offsets"]
    assert eng.TermCount() == len(terms)
    for i, t in enumerate(terms):
        docs, tfs = eng.decode_list(str(t))
        assert np.array_equal(docs, z["docs"][offs[i]:offs[i + 1]]), t
        assert np.array_equal(tfs, z["tfs"][offs[i]:offs[i + 1]]), t
    info = eng.info()
    checksum, _ = eng.decode_all()
    assert checksum == int(z["docs"].astype(np.uint64).sum() + z["tfs"].astype(np.uint64).sum())
    assert info.n_postings == len(z["docs"])


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("k,fn", [(10, "ref_top10.txt.gz"), (3, "ref_top3.txt.gz")])
def test_topk_vs_reference(engines, name, k, fn):
    from wiser_b200 import SearchQuery
    eng, _, d = engines[name]
    ref = read_ref_results(os.path.join(d, fn))
    full = read_ref_results(os.path.join(d, "ref_full.txt.gz"))
    qs = [SearchQuery(*parse_query_line(l), n_results=k) for l in open(os.path.join(d, "queries.txt"))]
    res = eng.SearchBatch(qs)
    assert len(res) == len(ref)
    for q, r, (rd, rs, rdf), (fd, fs, _) in zip(qs, res, ref, full):
        assert r.doc_freqs == rdf, q.terms
        check_topk(rd, rs, [e.doc_id for e in r.entries], [e.doc_score for e in r.entries], fd, fs,
                   what=" ".join(q.terms))


@pytest.mark.parametrize("name", FIXTURES)
def test_full_intersection_vs_reference(engines, name):
    """k = 10^6 (collect path): bit-exact intersection doc-id sets and scores."""
    from wiser_b200 import SearchQuery
    eng, _, d = engines[name]
    full = read_ref_results(os.path.join(d, "ref_full.txt.gz"))
    lines = open(os.path.join(d, "queries.txt")).read().split("\n")[:-1]
    qs = [SearchQuery(*parse_query_line(l), n_results=1000000) for l in lines]
    # collect-mode segments are sized by the shortest list; keep batches modest
    for lo in range(0, len(qs), 512):
        chunk = qs[lo:lo + 512]
        hits_needed = max(1, max(min([eng.PostinglistSizes([t]).get(t, 0) for t in q.terms] or [0])
                                 for q in chunk))
        for q in chunk:
            q.n_results = max(hits_needed, 33)
        res = eng.SearchBatch(chunk)
        for q, r, (fd, fs, fdf) in zip(chunk, res, full[lo:lo + 512]):
            assert r.doc_freqs == fdf
            check_full(fd, fs, [e.doc_id for e in r.entries], [e.doc_score for e in r.entries],
                       what=" ".join(q.terms))


@pytest.mark.parametrize("name", ["bloom3", "wikibloom4"])
@pytest.mark.parametrize("factor", [0, 1, 10])
def test_bloom_enabled_indexes_vs_reference(golden_dir, name, factor):
    """Indexes carrying the reference's Bloom-begin/-end sections (tests_18.cc:283-359): phrase and
    plain results equal the reference's at every bloom_enable_factor (the GPU path verifies the
    positions themselves, so the factor cannot change anything), top-10 and full intersections."""
    from wiser_b200 import GpuVacuumEngine, SearchQuery
    d = os.path.join(golden_dir, name)
    eng = GpuVacuumEngine(d, bloom_enable_factor=factor).Load()
    ref = read_ref_results(os.path.join(d, f"ref_top10_f{factor}.txt.gz"))
    full = read_ref_results(os.path.join(d, "ref_full.txt.gz"))
    qs = [SearchQuery(*parse_query_line(l), n_results=10) for l in open(os.path.join(d, "queries.txt"))]
    res = eng.SearchBatch(qs)
    assert len(res) == len(ref)
    for q, r, (rd, rs, rdf), (fd, fs, _) in zip(qs, res, ref, full):
        assert r.doc_freqs == rdf, q.terms
        check_topk(rd, rs, [e.doc_id for e in r.entries], [e.doc_score for e in r.entries], fd, fs,
                   what=" ".join(q.terms))
    for q in qs:
        q.n_results = 100
    res = eng.SearchBatch(qs)
    for q, r, (fd, fs, fdf) in zip(qs, res, full):
        check_full(fd, fs, [e.doc_id for e in r.entries], [e.doc_score for e in r.entries], what=" ".join(q.terms))
    eng.close()


def test_single_query_api_and_known_answers(engines):
    from wiser_b200 import SearchQuery
    eng, _, _ = engines["hello3"]
    r = eng.Search(SearchQuery(["wisconsin"]))
    assert [e.doc_id for e in r.entries] == [1] and r.entries[0].doc_score == 1.0925692944940748
    r = eng.Search(SearchQuery(["hello", "world"]))
    assert [e.doc_id for e in r.entries] == [2, 0] and r.doc_freqs == [3, 2]
    assert r.entries[0].doc_score == 0.67743596792765004
    assert r.entries[1].doc_score == 0.67229217626054072
    r = eng.Search(SearchQuery(["hello"], n_results=0))
    assert r.Size() == 0 and r.doc_freqs == []
    r = eng.Search(SearchQuery(["hello", "nosuchterm"]))
    assert r.Size() == 0 and r.doc_freqs == []
    assert eng.PostinglistSizes(["hello", "world", "zzz"]) == {"hello": 3, "world": 2}
    assert eng.TermCount() == 4


@pytest.mark.parametrize("n_shards", [2, 3])
def test_document_partitioned_shards_merge(golden_dir, n_shards):
    """SURVEY §8e on ONE GPU: each shard is its own index; per-shard top-k lists merged by the
    cross-shard merge kernel must equal the unsharded result bit-for-bit."""
    import ctypes as C
    import torch
    from wiser_b200 import Batch, GpuVacuumEngine, SearchQuery
    from wiser_b200.capi import HIT_DTYPE, check, lib
    d = os.path.join(golden_dir, "zipf2k")
    lines = open(os.path.join(d, "queries.txt")).read().split("\n")[:-1]
    qs = [SearchQuery(*parse_query_line(l), n_results=10) for l in lines]
    whole = GpuVacuumEngine(d).Load()
    qarr = whole.make_queries(qs)
    ref_hits, ref_n, _, _ = whole.search_batch(qarr, 10)
    n = len(qs)
    gathered = torch.zeros((n_shards, n, 10, 16), dtype=torch.uint8, device="cuda")
    gathered_n = torch.zeros((n_shards, n), dtype=torch.int32, device="cuda")
    shards = []
    for s in range(n_shards):
        e = GpuVacuumEngine(d, shard=s, n_shards=n_shards).Load()
        b = Batch(e, qarr, 10)
        b.run()
        b.sync()
        h, nh = b.fetch()
        gathered[s] = torch.from_numpy(h.view(np.uint8).reshape(n, 10, 16)).cuda()
        gathered_n[s] = torch.from_numpy(nh).cuda()
        shards.append((e, b))
    out = torch.zeros((n, 10, 16), dtype=torch.uint8, device="cuda")
    out_n = torch.zeros(n, dtype=torch.int32, device="cuda")
    check(lib().wsr_merge_topk_device(gathered.data_ptr(), gathered_n.data_ptr(), n_shards, n, 10,
                                      out.data_ptr(), out_n.data_ptr(), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(n, 160).view(HIT_DTYPE).reshape(n, 10)
    got_n = out_n.cpu().numpy()
    assert np.array_equal(got_n, ref_n)
    for i in range(n):
        k = ref_n[i]
        assert np.array_equal(got["doc_id"][i, :k], ref_hits["doc_id"][i, :k]), lines[i]
        assert np.array_equal(got["score"][i, :k].view(np.uint64), ref_hits["score"][i, :k].view(np.uint64))


def test_partition_directories_with_global_stats(golden_dir):
    """SURVEY §8e deployment shape: every shard opens ITS OWN partition directory (indexed
    separately by the reference, local doc ids), receives the collection statistics, and the
    merged per-shard top-k equals the single-index result bit for bit."""
    import torch
    from wiser_b200 import Batch, GpuVacuumEngine, SearchQuery
    from wiser_b200.capi import HIT_DTYPE, check, lib
    d = os.path.join(golden_dir, "zipf2k")
    lines = open(os.path.join(d, "queries.txt")).read().split("\n")[:-1]
    qs = [SearchQuery(*parse_query_line(l), n_results=10) for l in lines]
    whole = GpuVacuumEngine(d).Load()
    ref_hits, ref_n, _, _ = whole.search_batch(whole.make_queries(qs), 10)
    ora = OracleIndex(d)
    gdf = dict((l.split()[0], int(l.split()[1])) for l in open(os.path.join(d, "terms.txt")))
    n = len(qs)
    gathered = torch.zeros((2, n, 10, 16), dtype=torch.uint8, device="cuda")
    gathered_n = torch.zeros((2, n), dtype=torch.int32, device="cuda")
    keep = []
    for s, base in enumerate([0, 1000]):
        e = GpuVacuumEngine(os.path.join(golden_dir, f"zipf2k_p{s}")).Load()
        terms = [e.term_at(i)[0] for i in range(e.TermCount())]
        e.set_global_stats(base, 2000, ora.avg_doc_len, np.array([gdf[t] for t in terms], np.uint32))
        b = Batch(e, e.make_queries(qs), 10)
        b.run()
        b.sync()
        h, nh = b.fetch()
        gathered[s] = torch.from_numpy(h.view(np.uint8).reshape(n, 10, 16)).cuda()
        gathered_n[s] = torch.from_numpy(nh).cuda()
        keep.append((e, b))
    out = torch.zeros((n, 10, 16), dtype=torch.uint8, device="cuda")
    out_n = torch.zeros(n, dtype=torch.int32, device="cuda")
    check(lib().wsr_merge_topk_device(gathered.data_ptr(), gathered_n.data_ptr(), 2, n, 10,
                                      out.data_ptr(), out_n.data_ptr(), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(n, 160).view(HIT_DTYPE).reshape(n, 10)
    got_n = out_n.cpu().numpy()
    assert np.array_equal(got_n, ref_n)
    for i in range(n):
        k = ref_n[i]
        assert np.array_equal(got["doc_id"][i, :k], ref_hits["doc_id"][i, :k]), lines[i]
        assert np.array_equal(got["score"][i, :k].view(np.uint64), ref_hits["score"][i, :k].view(np.uint64))


@pytest.mark.parametrize("merge_ratio_x4", [None, 1 << 30])
def test_generated_corpus_vs_oracle(tmp_path, monkeypatch, merge_ratio_x4):
    """A 200k-doc corpus from the native generator (multi-thousand-posting lists, skewed AND,
    3-5 term queries, phrases; doc ids beyond 2^16, so sparse lists have blocks spanning more than
    65536 docs: 32-bit record heads, 16-byte records, the probe's wide-block path): GPU top-10 vs
    the CPU oracle on the same directory, bit-exact scores."""
    import subprocess
    import gen_query_log
    from wiser_b200 import GpuVacuumEngine, SearchQuery
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = str(tmp_path / "c")
    subprocess.check_call([os.path.join(root, "wiser_b200", "wsr_gen_corpus"), "--out", d, "--docs", "200000",
                           "--vocab", "150000", "--seed", "11", "--positions", "1"], stdout=subprocess.DEVNULL)
    groups = gen_query_log.load_groups(os.path.join(d, "terms.txt"), 3000)
    lines = (gen_query_log.generate("two_term", groups, 400, 1) +
             gen_query_log.generate("two_term_hh", groups, 150, 2) +
             gen_query_log.generate("multi_term", groups, 150, 3) +
             gen_query_log.generate("single_high", groups, 60, 4) +
             gen_query_log.generate("single_low", groups, 60, 5) +
             gen_query_log.generate("two_term_ll", groups, 300, 8) +
             gen_query_log.generate("two_term_lh", groups, 200, 9) +
             gen_query_log.generate("phrase2", groups, 200, 6) +
             gen_query_log.generate("phrase3", groups, 80, 7))
    if merge_ratio_x4 is not None:     # every non-phrase two-term query down the merge path
        monkeypatch.setenv("WSR_MERGE_RATIO_X4", str(merge_ratio_x4))
    eng = GpuVacuumEngine(d).Load()
    ora = OracleIndex(d)
    qs = [SearchQuery(*parse_query_line(l), n_results=10) for l in lines]
    res = eng.SearchBatch(qs)
    for q, r in zip(qs, res):
        rd, rs, rdf = ora.search(q.terms, 10, is_phrase=q.is_phrase)
        fd, fs, _ = ora.search(q.terms, 1 << 30, is_phrase=q.is_phrase)
        assert r.doc_freqs == rdf
        check_topk(rd, rs, [e.doc_id for e in r.entries], [e.doc_score for e in r.entries], fd, fs,
                   what=" ".join(q.terms))
    # full intersections through the collect path for a sample (AND and phrase queries)
    sample = qs[:120:3] + qs[-280::7]
    for q in sample:
        q.n_results = 100000
    res = eng.SearchBatch(sample)
    for q, r in zip(sample, res):
        fd, fs, _ = ora.search(q.terms, 1 << 30, is_phrase=q.is_phrase)
        check_full(fd, fs, [e.doc_id for e in r.entries], [e.doc_score for e in r.entries], what=" ".join(q.terms))


def _same_hits(h1, n1, h2, n2, k):
    assert len(n2) == len(n1) and np.array_equal(n1, n2)
    mask = np.arange(k)[None, :] < n1[:, None]
    assert np.array_equal(h1["doc_id"][mask], h2["doc_id"][mask])
    assert np.array_equal(h1["score"][mask].view(np.uint64), h2["score"][mask].view(np.uint64))


def _host_planned(eng, q, k):
    """wsr_search_batch in chunks small enough (< 8192 queries) to be planned by the HOST planner."""
    hs, ns = [], []
    for lo in range(0, len(q), 4000):
        h, n, _, _ = eng.search_batch(q[lo:lo + 4000], k)
        hs.append(h)
        ns.append(n)
    if not hs:
        return eng.search_batch(q, k)[:2]
    return np.concatenate(hs), np.concatenate(ns)


def test_search_batch_device_planner_equals_host_planner(golden_dir):
    """Batches of >= 8192 queries (k <= 32) are planned on the GPU from the wsr_query array; the
    result must equal what the host planner gives for the same queries, and misuse must fail."""
    from wiser_b200 import GpuVacuumEngine
    d = os.path.join(golden_dir, "zipf2k")
    eng = GpuVacuumEngine(d).Load()
    base = open(os.path.join(d, "queries.txt"), "rb").read()
    q = eng.parse_query_log(base * 3, 10)          # ~14k queries incl. phrases, missing terms, 8-term queries
    assert len(q) >= 8192
    h1, n1 = _host_planned(eng, q, 10)
    h2, n2, _, _ = eng.search_batch(q, 10)
    _same_hits(h1, n1, h2, n2, 10)
    q3 = q.copy()
    q3["k"] = 3                                    # per-query k below the stride
    h1, n1 = _host_planned(eng, q3, 10)
    h2, n2, _, _ = eng.search_batch(q3, 10)
    _same_hits(h1, n1, h2, n2, 10)
    bad = q.copy()
    bad["k"][100] = 11                             # k > k_stride
    with pytest.raises(Exception):
        eng.search_batch(bad, 10)
    bad = q.copy()
    bad["term_ids"][200, 0] = 0x7fffffff           # term id outside the index
    bad["n_terms"][200] = 1
    with pytest.raises(Exception):
        eng.search_batch(bad, 10)


def test_search_log_device_front_end_equals_host_planner(golden_dir):
    """wsr_search_log parses, looks terms up and plans ON THE GPU (frontend.cu); it must return
    exactly what the host parser + host planner + wsr_search_batch return, for every line shape
    the reference's QueryProducerByLog accepts (query_pool.h:251-352)."""
    from wiser_b200 import GpuVacuumEngine
    d = os.path.join(golden_dir, "zipf2k")
    eng = GpuVacuumEngine(d).Load()
    base = open(os.path.join(d, "queries.txt"), "rb").read()
    odd = b"\n".join([b"", b"   ", b"t0  t1", b" t0 t1 ", b'"t0 t1"', b'"', b'""', b'" t0"', b"nosuchterm t0",
                      b"t0\r", b"\tt0 t3\t", b'"t1"', b"t1 t1", b'"t2 t2"', b"t0 t1 t2 t3 t4 t5 t6 t7"])
    for text in (base * 12 + odd + b"\n",        # > 64 KiB, ends with a newline
                 odd + b"\n" + base[:-1],        # last line without a newline
                 b"t0", b"\n", b"\n\n t1 \n",
                 # the line count is a kernel that reads 16 bytes per thread: lengths around it
                 b"t0\n" * 5, b"t0\n" * 5 + b"\n", b"t0\n" * 5 + b"t1", b"\n" * 31 + b"t3", b"\n" * 32,
                 b"t1 t0\n" * 5 + b"t2", b"t0\n" * 11):
        for k in (10, 1, 32):
            q = eng.parse_query_log(text, k)                       # host parser
            h1, n1 = _host_planned(eng, q, k)                      # host planner
            h2, n2 = eng.search_log(text, k)                       # device front end
            _same_hits(h1, n1, h2, n2, k)
    # sparse results: after one call whose results fill under 10 % of n*k, wsr_search_log packs the
    # hits on the GPU and scatters them on the host; dense logs go back to the direct copy
    sparse = b"nosuchterm t0\n" * 3000 + b"t0 t1\nt2\n" + b"t0 nosuchterm\n" * 2000 + b"t3 t1\n"
    q = eng.parse_query_log(sparse, 10)
    h1, n1, _, _ = eng.search_batch(q, 10)
    for _ in range(3):
        h2, n2 = eng.search_log(sparse, 10)
        _same_hits(h1, n1, h2, n2, 10)
    q = eng.parse_query_log(base, 10)
    h1, n1, _, _ = eng.search_batch(q, 10)
    for _ in range(2):
        h2, n2 = eng.search_log(base, 10)
        _same_hits(h1, n1, h2, n2, 10)
    # the same front end behind the device-resident batch API (multi-GPU path)
    from wiser_b200.engine import Batch
    q = eng.parse_query_log(base, 10)
    h1, n1, _, _ = eng.search_batch(q, 10)
    b = Batch(eng, q[:1], 10)
    assert b.reset_log(base, 10) == len(q)
    b.run()
    h2, n2 = b.fetch()
    _same_hits(h1, n1, h2, n2, 10)
    # k > 32 goes through the host planner (collect class) and still matches
    q = eng.parse_query_log(base, 40)
    h1, n1, _, _ = eng.search_batch(q, 40)
    h2, n2 = eng.search_log(base, 40)
    _same_hits(h1, n1, h2, n2, 40)
    # a line with more than 8 terms is refused by both front ends
    bad = b"t0 t1\n" + b" ".join(b"t%d" % i for i in range(9)) + b"\n"
    with pytest.raises(Exception):
        eng.parse_query_log(bad, 10)
    with pytest.raises(Exception):
        eng.search_log(bad, 10)


def test_counting_pass_equals_plain_run(golden_dir):
    """wsr_batch_count_work runs the counting instantiations of the kernels: same hits as
    wsr_batch_run, and the work counters are filled (they stay 0 after a plain run)."""
    from wiser_b200 import GpuVacuumEngine
    from wiser_b200.engine import Batch
    d = os.path.join(golden_dir, "zipf2k")
    eng = GpuVacuumEngine(d).Load()
    text = open(os.path.join(d, "queries.txt"), "rb").read()
    b = Batch(eng, eng.parse_query_log(text, 10), 10)
    b.run()
    b.sync()
    h1, n1 = b.fetch()
    plain = b.stats()
    assert plain.decoded_postings == 0 and plain.touched_bytes == 0 and plain.listed_postings > 0
    b.count_work()
    h2, n2 = b.fetch()
    st = b.stats()
    assert st.decoded_postings > 0 and st.touched_bytes > 0 and st.matches > 0
    assert np.array_equal(n1, n2)
    mask = np.arange(10)[None, :] < n1[:, None]
    assert np.array_equal(h1["doc_id"][mask], h2["doc_id"][mask])
    assert np.array_equal(h1["score"][mask].view(np.uint64), h2["score"][mask].view(np.uint64))


@pytest.mark.parametrize("name", ["wiki4", "zipf2k"])
@pytest.mark.parametrize("ratio_x4", [0, 4, 1 << 30])
def test_two_term_merge_and_probe_paths_vs_reference(golden_dir, name, ratio_x4, monkeypatch):
    """Both two-term paths against the reference files: WSR_MERGE_RATIO_X4 = 0 sends every two-term
    query down the probe path (filter + galloping), 2^30 sends every one down the merge path (both
    lists streamed once), 4 only exactly balanced lists. Small units (k = 3 and 10) and the
    per-unit threshold exchange of multi-unit queries are covered by the long lists of zipf2k."""
    from wiser_b200 import GpuVacuumEngine, SearchQuery
    monkeypatch.setenv("WSR_MERGE_RATIO_X4", str(ratio_x4))
    d = os.path.join(golden_dir, name)
    eng = GpuVacuumEngine(d).Load()
    full = read_ref_results(os.path.join(d, "ref_full.txt.gz"))
    lines = open(os.path.join(d, "queries.txt")).read().split("\n")[:-1]
    for k, fn in [(10, "ref_top10.txt.gz"), (3, "ref_top3.txt.gz")]:
        ref = read_ref_results(os.path.join(d, fn))
        qs, keep = [], []
        for i, l in enumerate(lines):
            terms, is_phrase = parse_query_line(l)
            if len(terms) == 2:
                qs.append(SearchQuery(terms, is_phrase, n_results=k))
                keep.append(i)
        assert len(qs) > 50
        res = eng.SearchBatch(qs)
        for q, r, i in zip(qs, res, keep):
            rd, rs, rdf = ref[i]
            fd, fs, _ = full[i]
            assert r.doc_freqs == rdf, q.terms
            check_topk(rd, rs, [e.doc_id for e in r.entries], [e.doc_score for e in r.entries], fd, fs,
                       what=f"ratio_x4={ratio_x4} k={k} " + " ".join(q.terms))
    eng.close()


def test_group_of_partitions_on_one_device(golden_dir):
    """wsr_group_* with two partition directories on ONE device (no NCCL involved): the library
    exchanges the collection statistics between the partitions itself, every partition parses and
    searches the log on the GPU, the per-partition top-k are merged on the device — and the result,
    doc_freqs included, equals the whole index's (and through it the reference's)."""
    from wiser_b200 import GpuVacuumEngine, SearchQuery
    from wiser_b200.capi import HIT_DTYPE, WSR_MAX_TERMS
    from wiser_b200.dist import ShardGroup
    d = os.path.join(golden_dir, "zipf2k")
    lines = [l for l in open(os.path.join(d, "queries.txt")).read().split("\n")[:-1] if not l.startswith('"')]
    text = ("\n".join(lines) + "\n").encode()
    n = len(lines)
    whole = GpuVacuumEngine(d).Load()
    ref_hits, ref_n = whole.search_log(text, 10)
    qs = [SearchQuery(*parse_query_line(l), n_results=10) for l in lines]
    ref_res = whole.SearchBatch(qs)
    group = ShardGroup([os.path.join(golden_dir, "zipf2k_p0"), os.path.join(golden_dir, "zipf2k_p1")], [0])
    hits = np.zeros((n + 1, 10), HIT_DTYPE)
    nh = np.zeros(n + 1, np.int32)
    dfs = np.zeros((n + 1, WSR_MAX_TERMS), np.uint32)
    ndf = np.zeros(n + 1, np.int32)
    assert group.search_log(text, 10, hits, nh, dfs, ndf) == n
    assert np.array_equal(nh[:n], ref_n)
    ora = OracleIndex(d)
    for i in range(n):
        k = ref_n[i]
        # the stored average length of the whole index is a running mean, the group's a weighted
        # mean of the partitions': scores agree to the last ulps, not bit for bit
        assert np.array_equal(hits["doc_id"][i, :k], ref_hits["doc_id"][i, :k]), lines[i]
        assert np.allclose(hits["score"][i, :k], ref_hits["score"][i, :k], rtol=1e-12, atol=0), lines[i]
        if ref_res[i].doc_freqs:
            assert list(dfs[i, :ndf[i]]) == ref_res[i].doc_freqs, lines[i]
    # device-resident form: plan once, run, fetch
    assert group.load_log(text, 10) == n
    group.run()
    h2, n2 = group.fetch()
    assert np.array_equal(n2, nh[:n])
    m = np.arange(10)[None, :] < n2[:, None]
    assert np.array_equal(h2["doc_id"][m], hits[:n]["doc_id"][m])
    assert np.array_equal(h2["score"][m].view(np.uint64), hits[:n]["score"][m].view(np.uint64))
    group.close()
    whole.close()


def test_search_log_doc_freqs(engines):
    """wsr_search_log_ex: the log path returns SearchResult::doc_freqs like Search() does."""
    from wiser_b200 import SearchQuery
    from wiser_b200.capi import HIT_DTYPE, WSR_MAX_TERMS
    eng, _, d = engines["zipf2k"]
    lines = open(os.path.join(d, "queries.txt")).read().split("\n")[:-1]
    lines += ["nosuchterm t1", "", "t1 t1"]
    text = ("\n".join(lines) + "\n").encode()
    n = len(lines)
    hits = np.zeros((n + 1, 10), HIT_DTYPE)
    nh = np.zeros(n + 1, np.int32)
    dfs = np.zeros((n + 1, WSR_MAX_TERMS), np.uint32)
    ndf = np.zeros(n + 1, np.int32)
    h, c = eng.search_log(text, 10, hits, nh, dfs, ndf)
    assert len(c) == n
    res = eng.SearchBatch([SearchQuery(*parse_query_line(l), n_results=10) for l in lines])
    for i, r in enumerate(res):
        assert list(dfs[i, :ndf[i]]) == r.doc_freqs, lines[i]
        assert int(nh[i]) == len(r.entries)


def test_search_log_zero_copy_equals_staged_results(engines):
    """wsr_search_log_ex with pinned result buffers (the kernels write rows, counts and doc_freqs
    straight into host memory) against the same call with pageable buffers (pinned stage + copy)."""
    from wiser_b200.capi import HIT_DTYPE, WSR_MAX_TERMS, PinnedArray
    eng, _, d = engines["zipf2k"]
    text = open(os.path.join(d, "queries.txt"), "rb").read()
    n = text.count(b"\n")
    for k in (10, 3):
        hp, cp = PinnedArray((n + 2, k), HIT_DTYPE), PinnedArray((n + 2,), np.int32)
        dp, mp = PinnedArray((n + 2, WSR_MAX_TERMS), np.uint32), PinnedArray((n + 2,), np.int32)
        hp.array["doc_id"][:] = -7     # stale contents must not leak into counted rows
        cp.array[:] = 99
        h1, c1 = eng.search_log(text, k, hp.array, cp.array, dp.array, mp.array)
        hs, cs = np.zeros((n + 2, k), HIT_DTYPE), np.zeros(n + 2, np.int32)
        ds, ms = np.zeros((n + 2, WSR_MAX_TERMS), np.uint32), np.zeros(n + 2, np.int32)
        h2, c2 = eng.search_log(text, k, hs, cs, ds, ms)
        assert len(c1) == len(c2) == n
        _same_hits(h2, c2, h1, c1, k)
        assert np.array_equal(mp.array[:n], ms[:n])
        m = np.arange(WSR_MAX_TERMS)[None, :] < ms[:n, None]
        assert np.array_equal(dp.array[:n][m], ds[:n][m])
