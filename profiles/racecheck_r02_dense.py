#!/usr/bin/env python
"""racecheck target: the barrier-synchronised round-2 kernels only (doubling method at 5x5 / 3x3, TMA-staged inverse, L2 inverse);
the eigensolver's flag-synchronised relay is covered by the litmus tests instead (tests/test_parity.py::test_qr_flag_protocol_*)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wk  # noqa: E402
from khepri_b200 import Engine  # noqa: E402

eng = Engine(workspace_cap_bytes=2 << 30)
st, srcs = wk.case_bzi((3, 3), 2, 2)
cl = wk.build_crystal(st, eng, method="doubling")
R, T = cl.solve_batch([s["wavelength"] for s in srcs], kps=[s["kp"] for s in srcs])
print("doubling R+T-1", float(np.abs(R + T - 1).max()))
st, srcs = wk.case_suh03()
cl = wk.build_crystal(st, eng, method="doubling")
R, T = cl.solve_batch([s["wavelength"] for s in srcs[:2]], te=1.0, tm=0.0)
print("suh03", R, T)
rng = np.random.default_rng(0)
for n in (50, 98, 130, 242, 300):              # 130 / 242: cluster of 4 / 8 CTAs; 300: blocked, 32-pivot block columns
    A = rng.standard_normal((1, n, n)) + 1j * rng.standard_normal((1, n, n))
    print("zinv", n, float(np.abs(eng.zinv(A).cpu().numpy() @ A - np.eye(n)).max()))
