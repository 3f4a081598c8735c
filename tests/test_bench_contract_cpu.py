"""The reference arm of bench.py (`--impl reference`) runs without a GPU: it times the reference
engine (oracle/_ref/ref_tool when built, else the CPU oracle port) on a bounded sample of the
same workload. Checked here at toy scale: the JSON line carries every key of the contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--docs", "20000", "--vocab", "20000", "--queries", "400",
                          "--high-df", "200", "--cpu-sample", "400", "--dir", str(tmp_path)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.split("\n") if l.startswith("{")]
    assert len(lines) == 1                      # exactly one JSON line on stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["metric"] == "bm25_and_top10_listed_postings_per_s" and j["unit"] == "postings/s"
    assert j["steps"] == 2 and j["warmup"] == 1 and j["higher_is_better"] is True
    assert j["value"] > 0 and j["ms_per_step"] > 0
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


def test_reference_arm_replays_every_partition_at_n2(tmp_path):
    """At N > 1 the reference arm lists what our arm lists: rank 0 replays every one of the N
    partition directories with the same log (the other ranks exit without work)."""
    base = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
            "--warmup", "1", "--docs", "20000", "--vocab", "20000", "--queries", "400", "--high-df", "200",
            "--cpu-sample", "400", "--dir", str(tmp_path)]
    env = dict(os.environ, WORLD_SIZE="2", RANK="1", LOCAL_RANK="1")
    idle = subprocess.run(base, capture_output=True, text=True, timeout=600, env=env)
    assert idle.returncode == 0 and not [l for l in idle.stdout.split("\n") if l.startswith("{")]
    env = dict(os.environ, WORLD_SIZE="2", RANK="0", LOCAL_RANK="0")
    out = subprocess.run(base, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    j = json.loads([l for l in out.stdout.split("\n") if l.startswith("{")][0])
    assert j["n_gpus"] == 2 and "x 2 GPU(s)" in j["config"]["workload"]
    assert j["config"]["partitioning"].startswith("document-partitioned x2")
    one = subprocess.run(base[:4] + ["--gpus", "1"] + base[6:], capture_output=True, text=True, timeout=600)
    j1 = json.loads([l for l in one.stdout.split("\n") if l.startswith("{")][0])
    # two partitions list about twice the postings of one and take about twice as long
    assert os.path.isdir(os.path.join(str(tmp_path), "c_d20000_v20000_mu5.34_s1_p1of2"))
    assert j["ms_per_step"] > 0 and j1["n_gpus"] == 1
