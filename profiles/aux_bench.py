#!/usr/bin/env python
"""Secondary measurements for the other BASELINE configs and the HBM-bound kernels (one JSON line each).
Run on the GPU box:  python profiles/aux_bench.py > gpurun_out/aux_bench.jsonl"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from khepri_b200 import Crystal, Engine, Expansion, Layer  # noqa: E402
from tests import cases  # noqa: E402
from tests.util import build_crystal  # noqa: E402

HBM_PEAK = 6540.5        # GB/s, MEASURED_PEAKS.json (copy bandwidth measured on this pool's B200)
eng = Engine(workspace_cap_bytes=48 << 30)


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def emit(**kw):
    print(json.dumps(kw), flush=True)


# ---- convolution matrix: pruned DFT + gather; algorithmic bytes = Nx*Ny*8 read + N^2*16 written per layer
for L, res, pw in ((64, 512, (15, 15)), (16, 2048, (15, 15)), (256, 128, (7, 7))):
    pix = torch.rand((L, res, res), dtype=torch.float64, device="cuda")
    ms_wall, _ = timed(lambda: eng.convmat(pix, pw), reps=5, warm=2)
    # kernel time of the three launches (CUDA events around each, kh_profile): the Python call adds allocations of the outputs
    import ctypes as C
    eng.lib.kh_profile_begin()
    for _ in range(5):
        eng.convmat(pix, pw)
    buf = C.create_string_buffer(1 << 16); eng.lib.kh_profile_end(buf, len(buf))
    ms = sum(float(l.split()[2]) for l in buf.value.decode().splitlines() if l.split() and l.split()[0] in ("dft1", "dft2", "gather")) / 5
    N = pw[0] * pw[1]
    gb = L * (res * res * 8 + N * N * 16) / 1e9
    emit(kernel="convmat (dft1+dft2+gather)", layers=L, pixmap=[res, res], pw=list(pw), ms=ms, ms_python_call=ms_wall, algorithmic_GB=gb,
         achieved_GBps=gb / (ms * 1e-3), hbm_peak_GBps=HBM_PEAK, frac=gb / (ms * 1e-3) / HBM_PEAK)

# ---- sweeps of the other configs (solves/s, device-resident inputs)
def sweep_rate(name, st, wl, kp, pol, reps=3):
    cl = build_crystal(st, eng)
    plan = cl._get_plan(False)
    w, k, p = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (wl, kp.astype(complex), pol.astype(complex)))
    ms, res = timed(lambda: eng.solve_batch(plan, w, k, p, want_flux=True), reps=reps, warm=1)
    rt = res["RT"].cpu().numpy()
    emit(config=name, harmonics=list(st["pw"]), n=2 * st["pw"][0] * st["pw"][1], solves=len(wl), ms=ms, solves_per_s=len(wl) / (ms * 1e-3),
         finite=bool(np.isfinite(rt).all()), max_abs_R_plus_T_minus_1=float(np.nanmax(np.abs(rt.sum(1) - 1))))

freqs = np.linspace(0.49, 0.6, 151)
kx = np.linspace(0, 0.3 * np.pi, 64)
wl = np.tile(1 / freqs, 64); kp = np.stack([np.repeat(kx, 151), np.zeros(151 * 64)], 1); pol = np.tile([[1.0, 0.0]], (wl.size, 1))
sweep_rate("C1 suh03 5x5 (151 freqs x 64 kx)", cases.holey_pair(5, 128), wl, kp, pol)
st = cases.holey_pair(9, 128)
sweep_rate("holey pair 9x9 (151 freqs x 8 kx)", st, wl[:151 * 8], kp[:151 * 8], pol[:151 * 8])
fr = np.linspace(0.4 / 1.414, 0.65 / 1.414, 200); kxw = np.linspace(0, 0.99 * np.pi, 200)[:4]
wlw = np.tile(1 / fr, 4); kpw = np.stack([np.repeat(kxw, 200), np.zeros(800)], 1); polw = np.ones((800, 2))
sweep_rate("C3 woodpile 11x11 (200 freqs x 4 kx)", cases.woodpile_structure((11, 11)), wlw, kpw, polw, reps=2)

# ---- 13x13 and 15x15 direct bases (the large end of the north-star range; Twisted.ipynb cell 6-8 geometry: two pixmap layers)
def two_layer_structure(pp, res=256):
    pm = cases.disc_pixmap((res, res), 4.0, (0.0, 0.0), 0.25, 1.0)
    layers = {"A": ("pixmap", pm, 0.2), "B": ("pixmap", pm.T.copy() + 0.5 * cases.rect_pixmap((res, res), 0.0, (0.1, 0.0), (0.3, 0.5), 1.0), 0.2), "U": ("uniform", 1.0, 0.3)}
    return cases._st((pp, pp), layers, ["A", "U", "B"])

for pp, nf in ((13, 120), (15, 60)):
    fq = np.linspace(0.7, 0.83, nf)
    sweep_rate(f"direct {pp}x{pp} supercell basis, two pixmap layers ({nf} freqs)", two_layer_structure(pp), 1 / fq, np.zeros((nf, 2)), np.tile([[1.0, 0.0]], (nf, 1)), reps=2)

# ---- C4: twisted bilayer (3,3)+(3,3), n = 162 (extended RCWA)
tw = cases.twisted_case()
e1, e2 = Expansion(tw["pw"]), Expansion(tw["pw"])
e1.rotate(0.2); e2.rotate(-0.2)
cl = Crystal.from_expansion(e1 + e2, engine=eng)
cl.add_layer("upper", Layer.pixmap(e1, tw["pixmap"], 0.2), extended=True)
cl.add_layer("lower", Layer.pixmap(e2, tw["pixmap"], 0.2), extended=True)
cl.add_layer("inter", Layer.uniform(e1, 1, 0.3), extended=True)
cl.set_device(["upper", "inter", "lower"])
fq = np.linspace(0.7, 0.83, 50)
t0 = time.perf_counter(); R, T = cl.solve_batch(1 / fq, te=1, tm=0); torch.cuda.synchronize(); t1 = time.perf_counter()
t0 = time.perf_counter(); R, T = cl.solve_batch(1 / fq, te=1, tm=0); torch.cuda.synchronize(); t1 = time.perf_counter()
emit(config="C4 twisted bilayer (3,3)+(3,3) n=162, 50 freqs (e2e, host buffers)", ms=(t1 - t0) * 1e3, solves_per_s=50 / (t1 - t0),
     max_abs_R_plus_T_minus_1=float(np.abs(R + T - 1).max()))

# ---- C5: field maps 9x9, 256x256x128 volume (sliced holey pair), per frequency
st, src, _ = cases.case_fields(9, slices=4, res=128)
x = np.linspace(0, 1, 256); y = np.linspace(0, 1, 256); z = np.linspace(0.0001, 2.2, 128)
X, Y = np.meshgrid(x, y, indexing="xy")
cl = build_crystal(st, eng, fields=True)
cl.set_source(**src)
def one_volume():
    cl.solve()
    inc = np.hstack(cl.get_source_as_field_vectors())
    return cl._fields_points(X, Y, [float(v) for v in z], inc)
ms, F = timed(one_volume, reps=2, warm=1)
out_gb = F.numel() * 16 / 1e9
emit(config="C5 fields_volume 9x9, 256x256x128, one frequency (solve with retained eigenspaces + fields, device output)", ms=ms,
     output_GB=out_gb, achieved_GBps=out_gb / (ms * 1e-3), hbm_peak_GBps=HBM_PEAK, frac=out_gb / (ms * 1e-3) / HBM_PEAK,
     volumes_per_s=1e3 / ms, finite=bool(torch.isfinite(torch.view_as_real(F)).all().item()))
