"""Run under compute-sanitizer --tool racecheck on the GPU box: small eigenproblems + one solve."""
import numpy as np
from khepri_b200 import Engine
eng = Engine()
rng = np.random.default_rng(0)
for n in (18, 50):
    A = rng.standard_normal((2, n, n)) + 1j * rng.standard_normal((2, n, n))
    w, W, info = eng.zgeev(A)
    W = W.cpu().numpy(); w = w.cpu().numpy()
    print(n, "resid", np.abs(A @ W - W * w[:, None, :]).max(), info.cpu().numpy())
    Ai = eng.zinv(A).cpu().numpy()
    print(n, "inv", np.abs(Ai @ A - np.eye(n)).max())
