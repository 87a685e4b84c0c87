"""Developer tool: per-phase cycle counts of the QR kernel (library built with -DKH_QR_TIMING)."""
import ctypes as C, numpy as np, sys
sys.path.insert(0, ".")
from khepri_b200 import Engine
from tests import cases
from tests.util import build_crystal
eng = Engine(lib_path=__import__("os").environ.get("KH_TLIB", "tests/hostemu/libkh_timing.so"), device="cuda")
st = cases.bzi_structure((7, 7))
cl = build_crystal(st, eng)
wl = 1 / np.linspace(0.8, 1.0, 8)
out = (C.c_longlong * 16)()
eng.lib.kh_qr_timing.argtypes = [C.POINTER(C.c_longlong)]
eng.lib.kh_qr_timing(out)            # clear
R, T = cl.solve_batch(wl, kps=np.tile([[0.3, 0.2]], (8, 1)), te=1.0, tm=1.0)
eng.lib.kh_qr_timing(out)
names = ["total", "scan", "shift", "sweep", "delayed", "sweeps", "rotations"]
print({k: int(v) for k, v in zip(names, out)})
print("zhessz phases (balance+norm+v, matvec, update, out+shift, z-dots, z-update):", [int(out[i]) for i in range(8, 14)])
print("cycles/rotation in sweep:", out[3] / max(1, out[6]), " cycles/sweep fixed (scan+shift):", (out[1] + out[2]) / max(1, out[5]))
