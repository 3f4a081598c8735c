"""GPU tests of the C++ host side: the replay driver (wsr_replay) and, through it, the
GpuVacuumEngine adapter — batch mode and the multi-threaded Search() mode whose concurrent
callers are coalesced into GPU batches — against the committed reference results."""
import os
import subprocess

import pytest

from oracle_py import read_ref_results
from parity import check_topk

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPLAY = os.path.join(ROOT, "wiser_b200", "wsr_replay")


@pytest.mark.parametrize("mode,extra", [("batchlog", ["-batch_size=1000"]), ("locallog", ["-n_threads=16"])])
def test_replay_driver_matches_reference(golden_dir, tmp_path, mode, extra):
    d = os.path.join(golden_dir, "zipf2k")
    out = str(tmp_path / "dump.txt")
    log = subprocess.check_output([REPLAY, f"-engine=gpu:vacuum_dump:{d}", f"-query_path={d}/queries.txt",
                                   "-n_results=10", f"-exp_mode={mode}", f"-dump={out}"] + extra).decode()
    assert "WSR_REPLAY_JSON" in log
    got = read_ref_results(out)
    ref = read_ref_results(os.path.join(d, "ref_top10.txt.gz"))
    full = read_ref_results(os.path.join(d, "ref_full.txt.gz"))
    lines = open(os.path.join(d, "queries.txt")).read().split("\n")[:-1]
    assert len(got) == len(ref) == len(lines)
    for line, (gd, gs, gdf), (rd, rs, rdf), (fd, fs, _) in zip(lines, got, ref, full):
        assert gdf == rdf, line
        check_topk(rd, rs, gd, gs, fd, fs, what=line)


def test_replay_batchlog_timing_passes_use_the_device_front_end(golden_dir, tmp_path):
    """-repeat=3 with -dump: pass 0 goes through wsr_search_batch (it needs doc_freqs), passes 1-2
    send the log text through wsr_search_log (GPU front end); all passes must find the same hits."""
    import json
    d = os.path.join(golden_dir, "zipf2k")
    out = str(tmp_path / "dump.txt")
    log = subprocess.check_output([REPLAY, f"-engine=gpu:vacuum_dump:{d}", f"-query_path={d}/queries.txt",
                                   "-n_results=10", "-exp_mode=batchlog", "-batch_size=700", "-repeat=3",
                                   f"-dump={out}"]).decode()
    js = json.loads([l for l in log.split("\n") if l.startswith("WSR_REPLAY_JSON")][0][len("WSR_REPLAY_JSON"):])
    got = read_ref_results(out)
    one_pass = sum(len(g[0]) for g in got)
    assert js["queries"] == 3 * len(got)
    assert js["result_entries"] == 3 * one_pass
    assert js["listed_postings"] == 3 * sum(sum(g[2]) for g in got)


def test_replay_driver_over_a_partitioned_group(golden_dir, tmp_path):
    """wsr_replay -dirs=<p0>,<p1>: the two partition directories of zipf2k served as one group on
    one GPU (wsr_group_search_log). Doc ids, doc_freqs and result counts equal the reference's on
    the whole index; scores agree to the last ulps (the group's average document length is the
    weighted mean of the partitions', the whole index stores a running mean)."""
    import numpy as np
    d = os.path.join(golden_dir, "zipf2k")
    lines = [l for l in open(os.path.join(d, "queries.txt")).read().split("\n")[:-1]]
    keep = [i for i, l in enumerate(lines) if not l.startswith('"')]   # partitions were written without positions
    qpath = str(tmp_path / "q.txt")
    with open(qpath, "w") as f:
        f.write("\n".join(lines[i] for i in keep) + "\n")
    out = str(tmp_path / "dump.txt")
    dirs = ",".join(os.path.join(golden_dir, f"zipf2k_p{s}") for s in range(2))
    log = subprocess.check_output([REPLAY, f"-dirs={dirs}", "-devices=0", f"-query_path={qpath}", "-n_results=10",
                                   "-batch_size=900", f"-dump={out}"]).decode()
    assert "WSR_REPLAY_JSON" in log and '"mode": "group"' in log
    got = read_ref_results(out)
    ref = read_ref_results(os.path.join(d, "ref_top10.txt.gz"))
    full = read_ref_results(os.path.join(d, "ref_full.txt.gz"))
    assert len(got) == len(keep)
    for (gd, gs, gdf), i in zip(got, keep):
        rd, rs, rdf = ref[i]
        fd, fs, _ = full[i]
        assert len(gd) == len(rd), lines[i]
        if len(rd):
            assert gdf == rdf, lines[i]
        assert np.allclose(gs, rs, rtol=1e-12, atol=0), lines[i]
        # every returned doc is in the reference's full intersection with (to the last ulps) the
        # score the reference gives it; which of several near-equal docs makes the cut may differ
        ref_score = dict(zip(fd.tolist(), fs.tolist()))
        assert len(set(gd.tolist())) == len(gd)
        for doc, sc in zip(gd.tolist(), gs.tolist()):
            assert doc in ref_score and np.isclose(ref_score[doc], sc, rtol=1e-12, atol=0), lines[i]


def test_adapter_in_group_mode_serves_search_callers(golden_dir, tmp_path):
    """GpuVacuumEngine with GpuEngineOptions::partition_dirs (the reference maintainer's multi-GPU
    engine): 8 client threads call Search() on the two zipf2k partitions served as one collection;
    doc ids, counts and doc_freqs equal the reference's on the whole index."""
    import numpy as np
    d = os.path.join(golden_dir, "zipf2k")
    lines = [l for l in open(os.path.join(d, "queries.txt")).read().split("\n")[:-1]]
    keep = [i for i, l in enumerate(lines) if not l.startswith('"')][:1500]
    qpath = str(tmp_path / "q.txt")
    with open(qpath, "w") as f:
        f.write("\n".join(lines[i] for i in keep) + "\n")
    out = str(tmp_path / "dump.txt")
    dirs = ",".join(os.path.join(golden_dir, f"zipf2k_p{s}") for s in range(2))
    log = subprocess.check_output([REPLAY, f"-dirs={dirs}", "-devices=0", f"-query_path={qpath}", "-n_results=10",
                                   "-exp_mode=locallog", "-n_threads=8", f"-dump={out}"]).decode()
    assert "WSR_REPLAY_JSON" in log
    got = read_ref_results(out)
    ref = read_ref_results(os.path.join(d, "ref_top10.txt.gz"))
    full = read_ref_results(os.path.join(d, "ref_full.txt.gz"))
    assert len(got) == len(keep)
    for (gd, gs, gdf), i in zip(got, keep):
        rd, rs, rdf = ref[i]
        fd, fs, _ = full[i]
        assert len(gd) == len(rd) and gdf == rdf, lines[i]
        assert np.allclose(gs, rs, rtol=1e-12, atol=0), lines[i]
        ref_score = dict(zip(fd.tolist(), fs.tolist()))
        for doc, sc in zip(gd.tolist(), gs.tolist()):
            assert doc in ref_score and np.isclose(ref_score[doc], sc, rtol=1e-12, atol=0), lines[i]
