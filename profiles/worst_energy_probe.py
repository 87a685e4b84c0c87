#!/usr/bin/env python
"""Where does energy conservation of the whole configs[1] job deviate most, and do both layer methods agree there?"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wk
from khepri_b200 import Engine
sys.path.insert(0, ROOT)
import bench
eng = Engine(workspace_cap_bytes=60 << 30)
st, wl, kp, pol = bench.make_workload("bzi77-full", 0, 0, 4096)
out = {}
for m in ("doubling", "eig"):
    cl = wk.build_crystal(st, eng, method=m)
    R, T = cl.solve_batch(wl, kps=kp, te=pol[:, 0], tm=pol[:, 1])
    out[m] = np.stack([R, T], 1)
d = np.abs(out["doubling"].sum(1) - 1); e = np.abs(out["eig"].sum(1) - 1)
diff = np.abs(out["doubling"] - out["eig"]).max(1)
for name, arr in (("doubling |R+T-1|", d), ("eig |R+T-1|", e), ("|doubling - eig|", diff)):
    idx = np.argsort(arr)[-5:][::-1]
    print(json.dumps({"what": name, "worst": [{"i": int(i), "val": float(arr[i]), "wl": float(wl[i]), "kp": [float(kp[i, 0].real), float(kp[i, 1].real)],
                                              "RT_doubling": out["doubling"][i].tolist(), "RT_eig": out["eig"][i].tolist()} for i in idx],
                      "count_above_1e-9": int((arr > 1e-9).sum()), "count_above_1e-10": int((arr > 1e-10).sum())}))
