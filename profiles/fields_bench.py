#!/usr/bin/env python
"""C5 (BASELINE configs[4]): fields_volume 9x9 harmonics on a 256x256x128 grid, batched over frequencies.
Times the retained-eigenspace solve and the field reconstruction separately (CUDA events) and reports the
field kernels against the HBM roofline (output bytes 6*nz*nx*ny*16 per frequency).
Run on the GPU box:  python profiles/fields_bench.py [nfreq] > gpurun_out/fields_bench.jsonl"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from khepri_b200 import Engine  # noqa: E402
import workloads as cases  # noqa: E402
from workloads import build_crystal  # noqa: E402

HBM_PEAK = 6540.5
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 17
eng = Engine(workspace_cap_bytes=64 << 30)
st, src, _ = cases.case_fields(9, slices=4, res=128)
x = np.linspace(0, 1, 256); y = np.linspace(0, 1, 256); z = np.linspace(0.0001, 2.2, 128)
X, Y = np.meshgrid(x, y, indexing="xy")
cl = build_crystal(st, eng, fields=True)
cl.set_source(**src)
plan = cl._get_plan(True)
wl = 1 / np.linspace(0.49, 0.6, 51)[:nf]
kp = np.zeros((nf, 2), dtype=complex)
pol = np.tile([[1.0, 0.0]], (nf, 1)).astype(complex)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def inc_vectors():
    out = []
    for w in wl:
        cl.set_source(wavelength=float(w), te=1.0, tm=0.0)
        out.append(np.hstack(cl.get_source_as_field_vectors()))
    return np.asarray(out)


inc = inc_vectors()
cl.solve()            # sets stack_positions (crystal.py:196-203)
F = None
for rep in range(3):
    del F                            # (the 13 GB output of the previous repetition goes back to the caching allocator)
    torch.cuda.synchronize()
    e0 = ev()
    if rep == 2:
        eng.lib.kh_profile_begin()
    solved = eng.solve_batch(plan, wl, kp, pol, want_flux=True, want_fields=True)
    e1 = ev()
    if rep == 2:
        torch.cuda.synchronize()
        sbuf = C.create_string_buffer(1 << 16)
        eng.lib.kh_profile_end(sbuf, len(sbuf))
        e1 = ev()
    eng.lib.kh_profile_begin()
    F = eng.fields(plan, solved, wl, kp, inc, X.ravel(), Y.ravel(), z, cl.stack_positions, grid=(x, y))
    e2 = ev()
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    eng.lib.kh_profile_end(buf, len(buf))
kern, skern = {}, {}
for ln in buf.value.decode().strip().splitlines():
    nm, cnt, ms, work = ln.split()
    kern[nm] = dict(count=int(cnt), ms=round(float(ms), 3))
for ln in sbuf.value.decode().strip().splitlines():
    nm, cnt, ms, work = ln.split()
    skern[nm] = dict(count=int(cnt), ms=round(float(ms), 3))
out_gb = F.numel() * 16 / 1e9
ms_solve, ms_fields = e0.elapsed_time(e1), e1.elapsed_time(e2)
print(json.dumps({"config": f"C5 fields_volume 9x9, 256x256x128, {nf} frequencies batched", "ms_solve": ms_solve, "ms_fields": ms_fields,
                  "ms_fields_per_volume": ms_fields / nf, "output_GB": out_gb, "fields_GBps": out_gb / (ms_fields * 1e-3),
                  "hbm_peak_GBps": HBM_PEAK, "frac": out_gb / (ms_fields * 1e-3) / HBM_PEAK, "volumes_per_s_incl_solve": nf / ((ms_solve + ms_fields) * 1e-3),
                  "finite": bool(torch.isfinite(torch.view_as_real(F[0])).all().item()), "field_kernels_ms": kern, "solve_kernels_ms": skern}))
