#!/usr/bin/env python
"""kh_idft_batch (fourier.idft as a standalone operator) on device-resident inputs: one JSON line per shape.
Algorithmic bytes = M*npts*16 written (the phase matrix N*npts*16 is an internal temporary: written once, read once)."""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from khepri_b200 import Engine  # noqa: E402
from khepri_b200.engine import _ptr  # noqa: E402

HBM_PEAK = 6540.5
eng = Engine(); lib = eng.lib; dev = eng.device
rng = np.random.default_rng(3)
for M, N, npts in ((6, 81, 256 * 256), (6 * 128, 81, 256 * 256), (6, 225, 512 * 512), (1, 25, 128 * 128)):
    s = torch.from_numpy(rng.standard_normal((M, N)) + 1j * rng.standard_normal((M, N))).to(dev)
    kx = torch.from_numpy(rng.standard_normal(N) * 20 + 0j).to(dev); ky = torch.from_numpy(rng.standard_normal(N) * 20 + 0j).to(dev)
    x = torch.from_numpy(rng.random(npts)).to(dev); y = torch.from_numpy(rng.random(npts)).to(dev)
    out = torch.empty((M, npts), dtype=torch.complex128, device=dev)
    ws = eng.workspace(lib.kh_idft_work_bytes(N, npts))
    def run():
        rc = lib.kh_idft_batch(M, N, npts, _ptr(kx), _ptr(ky), _ptr(x), _ptr(y), _ptr(s), _ptr(out), _ptr(ws), ws.numel(), eng.stream())
        assert rc == 0
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ref = (s[0].cpu().numpy()[:, None] * np.exp(1j * (kx.cpu().numpy()[:, None] * x[:64].cpu().numpy() + ky.cpu().numpy()[:, None] * y[:64].cpu().numpy()))).sum(0)
    err = float(np.abs(out[0, :64].cpu().numpy() - ref).max())
    print(json.dumps({"kernel": "kh_idft_batch (idft_phase + zgemm)", "M": M, "N": N, "npts": npts, "ms": ms,
                      "out_GB": M * npts * 16 / 1e9, "phase_GB": N * npts * 16 / 1e9, "traffic_GBps_incl_phase": (M + 2 * N) * npts * 16 / 1e9 / (ms * 1e-3),
                      "hbm_peak_GBps": HBM_PEAK, "gemm_TFLOPs": 8.0 * M * N * npts / (ms * 1e-3) / 1e12, "max_abs_err_first64": err}), flush=True)
