#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/ref_tool, built from /root/reference by oracle/Makefile). Run in the build
container only (needs /root/reference); the outputs are committed and are what the CPU and
GPU test suites read — nothing at test time touches /root/reference.

Per fixture directory:
  my.tip my.vacuum my.doc_length   index files written by the reference's own dumper
  my.fdx my.fdt                    stub doc store (0 docs) so VacuumEngine::Load() still works
  terms.txt                        "term df" per line
  queries.txt                      query log (reference format)
  ref_top10.txt.gz                 reference Search() results, k=10   (ref_tool replay format)
  ref_top3.txt.gz                  same, k=3 (heap-eviction / tie behaviour)
  ref_full.txt.gz                  same, k=1e6 => full intersection with scores
  lists.npz                        every posting list decoded by the reference iterators
bloom3/ and wikibloom4/ are built WITH the reference's Bloom-begin/-end sections (ref_tool
buildbloom) and carry ref_top10_f{0,1,10}.txt.gz = results at bloom_enable_factor 0, 1, 10.
"""
import os
import shutil
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_query_log  # noqa: E402

REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_tool")
REF_TESTDATA = "/root/reference/src/qq_mem/src/testdata"
TMP = "/tmp/wsr_golden_tmp"


def run(*cmd):
    subprocess.check_call(list(cmd), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def write_stub_doc_store(d):
    # ChunkedDocStoreReader::LoadFdx (doc_store.h:365-392): varint n_doc_ids, varint buffer size
    with open(os.path.join(d, "my.fdx"), "wb") as f:
        f.write(bytes([0x00, 0x80, 0x80, 0x01]))
    with open(os.path.join(d, "my.fdt"), "wb") as f:
        f.write(b"\0")


def dump_lists_npz(index_dir, out_path):
    tmp = os.path.join(TMP, "lists.bin")
    run(REF_TOOL, "dumplists", index_dir, tmp)
    raw = open(tmp, "rb").read()
    terms, offs, docs, tfs = [], [0], [], []
    p = 0
    while p < len(raw):
        (ln,) = struct.unpack_from("<I", raw, p)
        terms.append(raw[p + 4:p + 4 + ln].decode())
        p += 4 + ln
        (df,) = struct.unpack_from("<I", raw, p)
        p += 4
        a = np.frombuffer(raw, np.uint32, 2 * df, p).reshape(df, 2)
        p += 8 * df
        docs.append(a[:, 0])
        tfs.append(a[:, 1])
        offs.append(offs[-1] + df)
    np.savez_compressed(out_path, terms=np.array(terms), offsets=np.array(offs, np.int64),
                        docs=np.concatenate(docs), tfs=np.concatenate(tfs))


def make_fixture(name, linedoc, queries, extra_ks=()):
    out = os.path.join(HERE, name)
    work = os.path.join(TMP, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    run(REF_TOOL, "build", linedoc, work)
    qpath = os.path.join(out, "queries.txt")
    with open(qpath, "w") as f:
        for q in queries:
            f.write(q + "\n")
    # reference results with the REAL doc store...
    run(REF_TOOL, "replay", work, qpath, "10", os.path.join(TMP, "real_top10.txt"))
    for fn in ("my.tip", "my.vacuum", "my.doc_length", "terms.txt"):
        shutil.copy(os.path.join(work, fn), os.path.join(out, fn))
    write_stub_doc_store(out)
    # ...must equal those with the stub doc store (return_snippets=false never reads it)
    for k, fn in [(10, "ref_top10.txt"), (3, "ref_top3.txt"), (1000000, "ref_full.txt")] + list(extra_ks):
        run(REF_TOOL, "replay", out, qpath, str(k), os.path.join(out, fn))
    a = open(os.path.join(TMP, "real_top10.txt")).read()
    b = open(os.path.join(out, "ref_top10.txt")).read()
    assert a == b, "stub doc store changed results"
    dump_lists_npz(out, os.path.join(out, "lists.npz"))
    import gzip
    for fn in sorted(os.listdir(out)):
        if fn.startswith("ref_") and fn.endswith(".txt"):
            with open(os.path.join(out, fn), "rb") as fi, \
                    gzip.GzipFile(os.path.join(out, fn + ".gz"), "wb", mtime=0) as fo:
                fo.write(fi.read())
            os.remove(os.path.join(out, fn))
    sz = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
    print(f"{name}: {len(queries)} queries, {sz / 1024:.0f} KiB")


def token_sequences(linedoc):
    """Per document the analysed token sequence, rebuilt from the tokenized + positions columns."""
    out = []
    for line in open(linedoc).read().split("\n")[1:]:
        cols = line.split("\t")
        if len(cols) < 5:
            continue
        terms = cols[2].split(" ")          # title, body, tokenized, offsets, positions, ...
        bags = cols[4].split(".")[:len(terms)]
        seq = {}
        for t, bag in zip(terms, bags):
            for p in bag.split(";"):
                if p:
                    seq[int(p)] = t
        out.append([seq[i] for i in sorted(seq)])
    return out


def make_bloom_fixture(name, linedoc, queries):
    """Index WITH the Bloom-begin/-end sections (ref_tool buildbloom = tests_18.cc:283-310). The
    reference is replayed with bloom_enable_factor 0, 1 and 10; the filter is a shortcut that must
    not change results, which is asserted here, and all three files are committed."""
    import gzip
    out = os.path.join(HERE, name)
    work = os.path.join(TMP, name)
    shutil.rmtree(out, ignore_errors=True)
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(out)
    run(REF_TOOL, "buildbloom", linedoc, work)
    qpath = os.path.join(out, "queries.txt")
    with open(qpath, "w") as f:
        for q in queries:
            f.write(q + "\n")
    for fn in ("my.tip", "my.vacuum", "my.doc_length"):
        shutil.copy(os.path.join(work, fn), os.path.join(out, fn))
    write_stub_doc_store(out)
    texts = {}
    for factor in (0, 1, 10):
        fn = os.path.join(out, f"ref_top10_f{factor}.txt")
        run(REF_TOOL, "replay", out, qpath, "10", fn, str(factor))
        texts[factor] = open(fn).read()
    assert texts[0] == texts[1] == texts[10], "bloom_enable_factor changed the reference's results"
    run(REF_TOOL, "replay", out, qpath, "1000000", os.path.join(out, "ref_full.txt"), "1")
    dump_lists_npz(out, os.path.join(out, "lists.npz"))
    z = np.load(os.path.join(out, "lists.npz"))
    with open(os.path.join(out, "terms.txt"), "w") as f:
        for t, a, b in zip(z["terms"], z["offsets"][:-1], z["offsets"][1:]):
            f.write(f"{t} {int(b - a)}\n")
    for fn in sorted(os.listdir(out)):
        if fn.startswith("ref_") and fn.endswith(".txt"):
            with open(os.path.join(out, fn), "rb") as fi, \
                    gzip.GzipFile(os.path.join(out, fn + ".gz"), "wb", mtime=0) as fo:
                fo.write(fi.read())
            os.remove(os.path.join(out, fn))
    sz = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
    print(f"{name}: {len(queries)} queries, {sz / 1024:.0f} KiB (bloom factors 0/1/10 agree)")


def bloom_fixtures():
    import random
    # F. the reference's bi-Bloom 3-doc fixture (a / a a b / a b c) and the phrases tests_18.cc
    #    expects to find ("a b", "b c") and not to find ("a x", "x c")
    make_bloom_fixture("bloom3", os.path.join(REF_TESTDATA, "iter_test_3_docs_tf_bi-bloom"),
                       ["a", "b", "c", "z", '"a b"', '"b c"', '"a x"', '"x c"', '"a c"', '"b a"', '"a a"',
                        '"a b c"', '"a a b"', '"b c a"', "a b", "b c", "a b c", '"a"', "a a"])
    # G. the 4-article Wikipedia fixture with prefix/suffix Bloom columns: n-grams that occur,
    #    shuffled ones that mostly do not, plain ANDs and single terms
    ld = os.path.join(REF_TESTDATA, "wiki_linedoc.toy.pre-suf-bloom")
    seqs = token_sequences(ld)
    rng = random.Random(31)
    qs = []
    for n, cnt in ((2, 300), (3, 120), (4, 40), (5, 15)):
        for _ in range(cnt):
            s = seqs[rng.randrange(len(seqs))]
            i = rng.randrange(len(s) - n)
            qs.append('"' + " ".join(s[i:i + n]) + '"')
    vocab = sorted({t for s in seqs for t in s})
    for n, cnt in ((2, 200), (3, 60)):
        for _ in range(cnt):
            qs.append('"' + " ".join(rng.choice(vocab) for _ in range(n)) + '"')
    qs += [" ".join(rng.sample(vocab, 2)) for _ in range(150)]
    qs += vocab[::7]
    make_bloom_fixture("wikibloom4", ld, qs)


def hello3_linedoc(path):
    # The 3-doc engine of the reference's tests.cc:407-421 ("hello world", "hello wisconsin",
    # "hello world big world"), written as WITH_POSITIONS linedoc.
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    docs = ["hello world", "hello wisconsin", "hello world big world"]
    with open(path, "w") as f:
        f.write("FIELDS_HEADER_INDICATOR###\tdoctitle\tbody\ttokenized\toffsets\tpositions\n")
        for i, body in enumerate(docs):
            toks = body.split()
            uniq, offs, poss, cur = [], {}, {}, 0
            for pos, t in enumerate(toks):
                if t not in offs:
                    uniq.append(t)
                    offs[t], poss[t] = [], []
                offs[t].append((cur, cur + len(t)))
                poss[t].append(pos)
                cur += len(t) + 1
            off_s = "".join("".join(f"{s},{e};" for s, e in offs[t]) + "." for t in uniq)
            pos_s = "".join("".join(f"{p};" for p in poss[t]) + "." for t in uniq)
            f.write(f"doc_{i}\t{body}\t{' '.join(uniq)}\t{off_s}\t{pos_s}\n")


def main():
    if not os.path.exists(REF_TOOL):
        sys.exit("oracle/_ref/ref_tool missing: run `make -C oracle ref` (needs /root/reference)")
    shutil.rmtree(TMP, ignore_errors=True)
    os.makedirs(TMP)
    if sys.argv[1:] == ["bloom"]:          # only the Bloom-enabled fixtures
        bloom_fixtures()
        return
    bloom_fixtures()

    # A. known-answer 3-doc engine
    ld = os.path.join(TMP, "hello3.linedoc")
    hello3_linedoc(ld)
    make_fixture("hello3", ld, ["wisconsin", "hello", "hello world", "world hello", "big",
                                "hello hello", "hello nosuchterm", "nosuchterm",
                                "hello world big", "big world hello wisconsin"])

    # B. the reference's own a / a b / a b c fixture
    make_fixture("abc3", os.path.join(REF_TESTDATA, "iter_test_3_docs"),
                 ["a", "b", "c", "d", "a b", "b a", "a b c", "c b a", "a a", "a d"])

    # C. the reference's 4-article Wikipedia fixture: every token of all-tokens.txt as a
    #    single-term query (the vacuum-vs-qq_mem differential of tests_15.cc:158-210), plus ANDs
    toks = open(os.path.join(REF_TESTDATA, "all-tokens.txt")).read().split()
    uniq = sorted(set(toks))
    import random
    rng = random.Random(11)
    qs = list(uniq)
    qs += [" ".join(rng.sample(uniq, 2)) for _ in range(300)]
    qs += [" ".join(rng.sample(uniq[:400], 3)) for _ in range(100)]
    qs += ["anarchist movement", "anarch movement polit", "the of and"]
    make_fixture("wiki4", os.path.join(REF_TESTDATA, "line_doc_with_positions"), qs)

    # D. seeded synthetic Zipf corpus, 2000 docs: multi-block lists, VInts tails, exact-128 cases
    ld = os.path.join(TMP, "zipf2k.linedoc")
    run(sys.executable, os.path.join(ROOT, "tools", "gen_linedoc.py"), "--docs", "2000",
        "--vocab", "3000", "--seed", "7", "--out", ld)
    work = os.path.join(TMP, "zipf2k_pre")
    run(REF_TOOL, "build", ld, work)
    groups = gen_query_log.load_groups(os.path.join(work, "terms.txt"), 200)
    qs = sorted(groups["low"] + groups["high"])          # every term once
    qs += gen_query_log.generate("two_term", groups, 600, 3)
    qs += gen_query_log.generate("multi_term", groups, 300, 4)
    qs += gen_query_log.generate("mix_aol", groups, 300, 5)
    qs += ["t0 t0", "t1 nosuch", "t5 t3 t5", "t0 t1 t2 t3 t4 t5 t6 t7", '"t0"', " t2  t1 "]
    # phrase queries (config 4): word n-grams that really occur, random n-grams that mostly do
    # not, repeated terms, a missing term
    import random
    prng = random.Random(21)
    bodies = [l.split("\t")[1].split(" ") for l in open(ld).read().split("\n")[1:] if l]
    for n, cnt in ((2, 260), (3, 100), (4, 30), (5, 10)):
        for _ in range(cnt):
            b = bodies[prng.randrange(len(bodies))]
            if len(b) > n:
                i = prng.randrange(len(b) - n)
                qs.append('"' + " ".join(b[i:i + n]) + '"')
    allt = sorted(groups["low"] + groups["high"])
    for n, cnt in ((2, 60), (3, 20)):
        for _ in range(cnt):
            qs.append('"' + " ".join(prng.choice(groups["high"]) for _ in range(n)) + '"')
    qs += ['"t0 t0"', '"t1 t0 t1"', '"t0 nosuch"', '"t3 t2"', '"t2 t3"', '"' + allt[5] + " " + allt[9] + '"']
    make_fixture("zipf2k", ld, qs)

    # E. the same corpus as two document partitions (docs 0-999 / 1000-1999), each indexed on
    #    its own by the reference: the per-partition directories of a document-partitioned
    #    deployment. Index files only; queries and expected results are zipf2k's.
    lines = open(ld).read().split("\n")
    header, docs = lines[0], [l for l in lines[1:] if l]
    for part, (lo, hi) in enumerate([(0, 1000), (1000, 2000)]):
        pld = os.path.join(TMP, f"zipf2k_p{part}.linedoc")
        with open(pld, "w") as f:
            f.write(header + "\n" + "\n".join(docs[lo:hi]) + "\n")
        work = os.path.join(TMP, f"zipf2k_p{part}")
        run(REF_TOOL, "build", pld, work)
        out = os.path.join(HERE, f"zipf2k_p{part}")
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        for fn in ("my.tip", "my.vacuum", "my.doc_length", "terms.txt"):
            shutil.copy(os.path.join(work, fn), os.path.join(out, fn))
        write_stub_doc_store(out)
        print(f"zipf2k_p{part}: docs {lo}-{hi - 1}")


if __name__ == "__main__":
    main()
