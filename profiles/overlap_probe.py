#!/usr/bin/env python
"""Probe: does splitting one step's batch over 2-3 CUDA streams (separate workspaces, same plan) overlap the latency-bound
eigensolver kernels of one part with the DMMA kernels of another?  bzi77, 4141 solves per step.  One JSON line per variant."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload
from tests.util import build_crystal
from khepri_b200.engine import Engine

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
engs = [Engine(workspace_cap_bytes=40 << 30) for _ in range(3)]
st, wl, kp, pol = make_workload("bzi77", 0, 0, 41)
cl = build_crystal(st, engs[0]); plan = cl._get_plan(False)
B = wl.size
ins = []
for s in range(8):
    _, w, k, p = make_workload("bzi77", 0, s, 41)
    ins.append((torch.from_numpy(w).to(dev), torch.from_numpy(k).to(dev), torch.from_numpy(p).to(dev)))
streams = [torch.cuda.Stream(dev) for _ in range(3)]

def run(parts, s):
    w, k, p = ins[s]
    if parts == 1:
        return [engs[0].solve_batch(plan, w, k, p, want_flux=True)["RT"]]
    cur = torch.cuda.current_stream(dev)
    edges = np.linspace(0, B, parts + 1).astype(int)
    outs = []
    for i in range(parts):
        streams[i].wait_stream(cur)
        with torch.cuda.stream(streams[i]):
            lo, hi = edges[i], edges[i + 1]
            outs.append(engs[i].solve_batch(plan, w[lo:hi], k[lo:hi], p[lo:hi], want_flux=True)["RT"])
    for i in range(parts):
        cur.wait_stream(streams[i])
    return outs

ref = None
for parts in (1, 2, 3, 1):
    for s in range(3):
        run(parts, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(3, 8):
        out = run(parts, s)
    e1.record(); torch.cuda.synchronize()
    rt = torch.cat(out).cpu().numpy()
    if ref is None: ref = rt
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"parts": parts, "ms_per_step": ms, "solves_per_s": B / ms * 1e3, "max_abs_diff_vs_1": float(np.abs(rt - ref).max())}), flush=True)
