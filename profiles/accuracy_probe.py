#!/usr/bin/env python
"""Accuracy of the two patterned-layer methods on full sweeps (run on the GPU box): energy conservation |R + T - 1| of
lossless stacks and the largest difference between the eigen-decomposition and the doubling method, per config.
    python profiles/accuracy_probe.py > gpurun_out/accuracy.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wk  # noqa: E402
from khepri_b200 import Engine  # noqa: E402

eng = Engine(workspace_cap_bytes=48 << 30)


def probe(name, st, wl, kp, pol, thetas=(6.0, 10.0, 14.0)):
    out = {}
    for method, th in [("eig", None)] + [("doubling", t) for t in thetas]:
        cl = wk.build_crystal(st, eng, method=method)
        if th is not None:
            eng.doubling_theta = th
        R, T = cl.solve_batch(wl, kps=kp, te=pol[:, 0], tm=pol[:, 1])
        out[(method, th)] = np.stack([R, T], 1)
    ref = out[("eig", None)]
    for (method, th), rt in out.items():
        d = np.abs(rt - ref).max(axis=1)
        i = int(np.argmax(d))
        print(json.dumps({"config": name, "method": method, "theta_slice": th, "max_abs_R_plus_T_minus_1": float(np.abs(rt.sum(1) - 1).max()),
                          "max_abs_diff_vs_eig": float(d.max()), "worst_source": {"index": i, "wavelength": float(wl[i]), "kp": [float(kp[i, 0].real), float(kp[i, 1].real)],
                                                                                 "RT": rt[i].tolist(), "RT_eig": ref[i].tolist()}}), flush=True)


freqs = np.linspace(0.49, 0.6, 151)
kx = np.linspace(0, 0.3 * np.pi, 64)
wl = np.tile(1 / freqs, 64); kp = np.stack([np.repeat(kx, 151), np.zeros(151 * 64)], 1).astype(complex); pol = np.tile([[1.0, 0.0]], (wl.size, 1)).astype(complex)
probe("C1 suh03 5x5 (151 freqs x 64 kx)", wk.holey_pair(5, 128), wl, kp, pol)
probe("holey pair 9x9 (151 freqs x 8 kx)", wk.holey_pair(9, 128), wl[:151 * 8], kp[:151 * 8], pol[:151 * 8])
st = wk.bzi_structure((7, 7))
kg = wk.bzi_kgrid((64, 64)).reshape(2, -1)
wls = 1 / np.linspace(0.8, 1.0, 101)
ks = kg[:, 1000:1041]
probe("C2 bzi 7x7 (41 k x 101 wl)", st, np.tile(wls, 41), np.repeat(ks.T, 101, axis=0).astype(complex), np.ones((4141, 2), dtype=complex))
fr = np.linspace(0.4 / 1.414, 0.65 / 1.414, 200); kxw = np.linspace(0, 0.99 * np.pi, 200)[:2]
probe("C3 woodpile 11x11 (200 freqs x 2 kx)", wk.woodpile_structure((11, 11)), np.tile(1 / fr, 2), np.stack([np.repeat(kxw, 200), np.zeros(400)], 1).astype(complex), np.ones((400, 2), dtype=complex))
