#!/usr/bin/env python
"""Host-side profile of the drop-in scalar loop (README.md:55-59: set_source; solve; poynting_flux_end per frequency)."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wk  # noqa: E402
from khepri_b200 import Engine  # noqa: E402

eng = Engine()
st, srcs = wk.case_suh03()
cl = wk.build_crystal(st, eng)


def loop():
    acc = 0.0
    for s in srcs:
        cl.set_source(**s)
        cl.solve()
        acc += sum(cl.poynting_flux_end())
    torch.cuda.synchronize()
    return acc


loop()
t0 = time.perf_counter(); loop(); dt = time.perf_counter() - t0
print(f"scalar loop: {len(srcs) / dt:.1f} solves/s ({dt / len(srcs) * 1e3:.3f} ms per solve), launches per solve {0}")
l0 = eng.lib.kh_launch_count(); loop(); print("launches per solve", (eng.lib.kh_launch_count() - l0) / len(srcs))
pr = cProfile.Profile(); pr.enable(); loop(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
