#!/usr/bin/env python
"""K1 alone: opens the C2 corpus bench.py generated and runs wsr_decode_all a few times (the
command ncu profiles for the decode kernel). Prints ms and GB/s of payload and of algorithmic bytes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wiser_b200 import GpuVacuumEngine  # noqa: E402

d = sys.argv[1] if len(sys.argv) > 1 else "/tmp/wsr_bench/c_d5000000_v5000000_mu5.34_s1_p0of1"
eng = GpuVacuumEngine(d, positions=False).Load()
info = eng.info()
ms = [eng.decode_all()[1] for _ in range(6)]
best = min(ms[1:])
print(f"decode_all ms {['%.3f' % m for m in ms]} best {best:.3f}: {info.n_postings / best / 1e6:.1f} G postings/s, "
      f"payload {info.payload_bytes / best / 1e6:.0f} GB/s, hbm_bytes {info.hbm_bytes / 1e9:.2f} GB")
eng.close()
