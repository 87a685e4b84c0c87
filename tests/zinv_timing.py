"""Developer tool: time the batched inverse (and check it) for a library variant: KH_TLIB=... python tests/zinv_timing.py [n] [batch]"""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from khepri_b200 import Engine
eng = Engine(lib_path=os.environ.get("KH_TLIB"), device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 98
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4141
rng = np.random.default_rng(1)
A = torch.from_numpy(rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))).cuda()
Ai = eng.zinv(A)
err = (torch.bmm(Ai[:8], A[:8]) - torch.eye(n, device="cuda", dtype=A.dtype)).abs().max().item()
torch.cuda.synchronize()
for _ in range(20):
    eng.zinv(A)
ms = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.zinv(A)
    e1.record(); torch.cuda.synchronize()
    ms = min(ms, e0.elapsed_time(e1) / 10)
print(f"kernel {os.environ.get('KH_ZINV_KERNEL', '0')}: n={n} batch={batch} {ms:.3f} ms  {8.0 * n**3 * batch / ms / 1e9:.2f} TFLOP/s  err {err:.2e}")
