"""Run under compute-sanitizer (--tool racecheck / memcheck) on the GPU box: the shared-memory DMMA inverse, the
unit-balanced / pipelined GEMM and the row-per-thread convolution-matrix stages on small ragged shapes."""
import numpy as np
import torch
from khepri_b200 import Engine
eng = Engine()
rng = np.random.default_rng(0)
for n in (16, 24, 50, 98, 104):
    A = rng.standard_normal((3, n, n)) + 1j * rng.standard_normal((3, n, n))
    A[1] = np.roll(A[1], 1, axis=0) * 1e-3 + np.eye(n)[::-1]
    Ai = eng.zinv(A).cpu().numpy()
    print(n, "inv", np.abs(Ai @ A - np.eye(n)).max())
for (M, N, K) in ((98, 98, 98), (50, 50, 50), (57, 9, 33), (130, 65, 17)):
    A = rng.standard_normal((3, M, K)) + 1j * rng.standard_normal((3, M, K))
    B = rng.standard_normal((3, K, N)) + 1j * rng.standard_normal((3, K, N))
    print((M, N, K), "gemm", np.abs(eng.zgemm(A, B).cpu().numpy() - A @ B).max())
for res, pw in ((64, (5, 5)), (96, (7, 3))):
    pix = torch.rand((2, res, res + 8), dtype=torch.float64, device="cuda")
    C = eng.convmat(pix, pw)
    print(res, pw, "convmat finite", bool(torch.isfinite(torch.view_as_real(C[0] if isinstance(C, tuple) else C)).all()))
