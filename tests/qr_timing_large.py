"""Developer tool: per-phase cycle counts of the tiled QR kernel (library built with -DKH_QR_TIMING), n beyond shared memory."""
import ctypes as C, numpy as np, sys, os
sys.path.insert(0, ".")
import torch
from khepri_b200 import Engine
eng = Engine(lib_path=os.environ.get("KH_TLIB", "tests/hostemu/libkh_timing.so"), device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 242
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 400
rng = np.random.default_rng(5)
A = rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))
out = (C.c_longlong * 16)()
eng.lib.kh_qr_timing.argtypes = [C.POINTER(C.c_longlong)]
eng.zgeev(A[:4])
eng.lib.kh_qr_timing(out)            # clear
eng.lib.kh_profile_begin()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); w, W, info = eng.zgeev(A); t1.record(); torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16); eng.lib.kh_profile_end(buf, len(buf))
print(buf.value.decode())
eng.lib.kh_qr_timing(out)
names = ["total", "scan", "shift", "sweep", "delayed", "sweeps", "rotations"]
print("n", n, "batch", batch, "ms", t0.elapsed_time(t1), "info", int(info.max()))
print({k: int(v) for k, v in zip(names, out)})
print("tiled sweep (tile load, chase, store+bulk):", [int(out[i]) for i in range(8, 11)])
print("cycles/rotation in sweep:", out[3] / max(1, out[6]), " cycles/sweep fixed (scan+shift):", (out[1] + out[2]) / max(1, out[5]), "delayed/sweep", out[4] / max(1, out[5]))
