#!/usr/bin/env python
"""Small invocations of every round-2 kernel for compute-sanitizer (memcheck / racecheck):
doubling method (series, even/odd conversion, self star products, collapsed runs, star_last3, zgemv2), TMA-staged inverse,
single-launch L2 inverse, eig method with retained eigenspaces + field maps, extended layer with fields."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wk  # noqa: E402
from khepri_b200 import Crystal, Engine, Expansion, Layer  # noqa: E402

eng = Engine(workspace_cap_bytes=4 << 30)
# doubling + eig, flux only (deep layer -> self star products; repeated layers -> collapsed runs)
st, srcs = wk.case_bzi((3, 3), 2, 2)
for m in ("doubling", "eig"):
    cl = wk.build_crystal(st, eng, method=m)
    eng.doubling_theta = 2.0 if m == "doubling" else 10.0
    R, T = cl.solve_batch([s["wavelength"] for s in srcs], kps=[s["kp"] for s in srcs])
    print(m, "R+T-1", float(np.abs(R + T - 1).max()))
eng.doubling_theta = 10.0
st, srcs = wk.case_suh03()
cl = wk.build_crystal(st, eng)
R, T, S = cl.solve_batch([s["wavelength"] for s in srcs[:3]], te=1.0, tm=0.0, return_S=True)
# inverses: smem resident (TMA staged), single-launch L2 variant, blocked
rng = np.random.default_rng(0)
for n in (18, 50, 98, 130, 242, 300):          # 130 / 242: one thread-block cluster of 4 / 8 CTAs per matrix
    A = rng.standard_normal((2, n, n)) + 1j * rng.standard_normal((2, n, n))
    Ai = eng.zinv(A).cpu().numpy()
    print("zinv", n, float(np.abs(Ai @ A - np.eye(n)).max()))
# fields (eig path, TMA-staged Hessenberg / replay) incl. an extended layer
st, src, (X, Y, z) = wk.case_fields(5)
cl = wk.build_crystal(st, eng, fields=True)
cl.set_source(**src); cl.solve(); E, H = cl.fields_volume(X, Y, z)
print("fields", float(np.abs(E).max()))
tw = wk.twisted_case()
e1, e2 = Expansion(tw["pw"]), Expansion(tw["pw"]); e1.rotate(0.1); e2.rotate(-0.1)
cl = Crystal.from_expansion(e1 + e2, engine=eng)
cl.add_layer("u", Layer.pixmap(e1, tw["pixmap"], 0.2), extended=True)
cl.add_layer("l", Layer.pixmap(e2, tw["pixmap"], 0.2), extended=True)
cl.add_layer("i", Layer.uniform(e1, 1, 0.3), extended=True)
cl.set_device(["u", "i", "l"], [True] * 3)
cl.set_source(wavelength=1.3, te=1, tm=0); cl.solve()
Xg, Yg, zg = wk.twisted_field_grid()
E, H = cl.fields_volume(Xg, Yg, zg)
print("twisted fields", float(np.abs(E).max()))
