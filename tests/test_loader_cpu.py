"""CPU tests of the PRODUCT's host side: the vacuum loader's HBM block layout (decoded by the
host restatement in tests/host_index_dump.cc) against the reference iterators' dumps, sharding,
and the native corpus generator (its vacuum files must read identically through the oracle)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "wiser_b200", "wsr_gen_corpus")


@pytest.fixture(scope="module")
def dump_tool(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bin") / "host_index_dump")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", out,
                           os.path.join(ROOT, "tests", "host_index_dump.cc"),
                           os.path.join(ROOT, "wiser_b200", "csrc", "host_index.cc")])
    return out


def read_dump(path):
    raw = open(path, "rb").read()
    out, p = {}, 0
    while p < len(raw):
        (ln,) = struct.unpack_from("<I", raw, p)
        term = raw[p + 4:p + 4 + ln].decode()
        p += 4 + ln
        (df,) = struct.unpack_from("<I", raw, p)
        p += 4
        a = np.frombuffer(raw, np.uint32, 2 * df, p).reshape(df, 2)
        p += 8 * df
        out[term] = (a[:, 0].copy(), a[:, 1].copy())
    return out


@pytest.mark.parametrize("name", ["hello3", "abc3", "wiki4", "zipf2k", "bloom3", "wikibloom4"])
def test_layout_decodes_to_reference_lists(golden_dir, dump_tool, tmp_path, name):
    d = os.path.join(golden_dir, name)
    out = str(tmp_path / "dump.bin")
    subprocess.check_call([dump_tool, d, "0", "1", out], stdout=subprocess.DEVNULL)
    got = read_dump(out)
    z = np.load(os.path.join(d, "lists.npz"))
    offs = z["offsets"]
    assert len(got) == len(z["terms"])
    for i, t in enumerate(z["terms"]):
        docs, tfs = got[str(t)]
        assert np.array_equal(docs, z["docs"][offs[i]:offs[i + 1]]), t
        assert np.array_equal(tfs, z["tfs"][offs[i]:offs[i + 1]]), t


@pytest.mark.parametrize("n_shards", [2, 3])
def test_shards_partition_every_list(golden_dir, dump_tool, tmp_path, n_shards):
    """Document partitioning: the shards' sub-lists concatenate to the whole list and respect the
    doc ranges [s*N/n, (s+1)*N/n)."""
    d = os.path.join(golden_dir, "zipf2k")
    z = np.load(os.path.join(d, "lists.npz"))
    offs = z["offsets"]
    parts = []
    for s in range(n_shards):
        out = str(tmp_path / f"s{s}.bin")
        info = subprocess.check_output([dump_tool, d, str(s), str(n_shards), out]).decode().split()
        lo, hi = int(info[-2]), int(info[-1])
        assert lo == 2000 * s // n_shards and hi == 2000 * (s + 1) // n_shards
        parts.append((read_dump(out), lo, hi))
    for i, t in enumerate(z["terms"]):
        docs = np.concatenate([p[0][str(t)][0] for p in parts])
        tfs = np.concatenate([p[0][str(t)][1] for p in parts])
        assert np.array_equal(docs, z["docs"][offs[i]:offs[i + 1]]), t
        assert np.array_equal(tfs, z["tfs"][offs[i]:offs[i + 1]]), t
        for got, lo, hi in parts:
            dd = got[str(t)][0]
            assert len(dd) == 0 or (dd.min() >= lo and dd.max() < hi)


@pytest.mark.skipif(not os.path.exists(GEN), reason="wsr_gen_corpus not built")
def test_native_corpus_generator_roundtrip(dump_tool, tmp_path):
    """The generator writes the reference's vacuum format: the CPU oracle (pinned to the
    reference) and the product loader must read the same postings from it; deterministic."""
    from oracle_py import OracleIndex
    d1, d2 = str(tmp_path / "c1"), str(tmp_path / "c2")
    for d in (d1, d2):
        subprocess.check_call([GEN, "--out", d, "--docs", "3000", "--vocab", "5000", "--seed", "5",
                               "--threads", "3" if d == d1 else "1"], stdout=subprocess.DEVNULL)
    for fn in ("my.vacuum", "my.tip", "my.doc_length"):
        assert open(os.path.join(d1, fn), "rb").read() == open(os.path.join(d2, fn), "rb").read(), fn
    out = str(tmp_path / "dump.bin")
    subprocess.check_call([dump_tool, d1, "0", "1", out], stdout=subprocess.DEVNULL)
    got = read_dump(out)
    ora = OracleIndex(d1)
    assert ora.num_docs == 3000 and ora.term_count == len(got)
    terms = ora.terms()
    for t in terms[:200] + terms[-200:] + terms[::37]:
        docs, tfs = ora.decode_list(t)
        assert np.array_equal(docs, got[t][0]) and np.array_equal(tfs, got[t][1]), t
        assert np.all(np.diff(docs.astype(np.int64)) > 0) and tfs.min() >= 1
    # df listed in terms.txt == posting-list sizes; exact-128 / multi-block lists exist
    dfs = dict((l.split()[0], int(l.split()[1])) for l in open(os.path.join(d1, "terms.txt")))
    assert all(len(got[t][0]) == n for t, n in dfs.items())
    assert max(dfs.values()) > 1000
