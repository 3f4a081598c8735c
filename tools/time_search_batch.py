"""Developer tool: times wsr_search_batch on the bench corpus (run bench.py once first so that the
corpus and the 100k two-term log exist under /tmp/wsr_bench). WSR_HOST_FRONTEND=1 forces the host planner."""
import sys, time, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wiser_b200 import GpuVacuumEngine
from wiser_b200.capi import HIT_DTYPE, PinnedArray
d = "/tmp/wsr_bench/c_d5000000_v5000000_mu5.34_s1_p0of1"
eng = GpuVacuumEngine(d, positions=False).Load()
text = open(d + "/q_two_term_n100000_h10000_s1.txt", "rb").read()
q = eng.parse_query_log(text, 10)
n = len(q)
hits = PinnedArray((n, 10), HIT_DTYPE); nh = PinnedArray((n,), np.int32)
for _ in range(3):
    eng.search_batch(q, 10, hits.array, nh.array)
t0 = time.perf_counter()
for _ in range(10):
    eng.search_batch(q, 10, hits.array, nh.array)
print("search_batch 100k queries: %.2f ms/call (WSR_HOST_FRONTEND=%s)" % ((time.perf_counter() - t0) * 100, os.environ.get("WSR_HOST_FRONTEND", "0")))
os._exit(0)
