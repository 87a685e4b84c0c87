"""Developer tool: time the batched complex GEMM per launcher variant (KH_ZGEMM_VARIANT is read once per process):
    python tests/zgemm_timing.py [n] [batch]            -> spawns one child per variant
Also prints the FP64 peaks (DFMA, DMMA, both interleaved)."""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, ".")

def child(n, batch):
    import numpy as np, torch
    from khepri_b200 import Engine
    eng = Engine(device="cuda")
    rng = np.random.default_rng(1)
    A = torch.from_numpy(rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))).cuda()
    B = torch.from_numpy(rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))).cuda()
    Cc = eng.zgemm(A, B)
    err = (Cc[:8] - torch.bmm(A[:8], B[:8])).abs().max().item()
    Ct = eng.zgemm(A, B, transA=True)
    errt = (Ct[:8] - torch.bmm(A[:8].transpose(1, 2), B[:8])).abs().max().item()
    torch.cuda.synchronize()
    for _ in range(100):            # clocks up
        eng.zgemm(A, B)
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            eng.zgemm(A, B)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 40)
    ms = best
    print(f"variant {os.environ.get('KH_ZGEMM_VARIANT', '0')}: n={n} batch={batch} {ms:.3f} ms  {8.0 * n**3 * batch / ms / 1e9:.2f} TFLOP/s  err {err:.2e} errT {errt:.2e}", flush=True)

def peaks():
    import torch
    from khepri_b200 import Engine
    eng = Engine(device="cuda")
    scratch = torch.zeros(16, dtype=torch.float64, device="cuda")
    for mode, name in ((0, "dfma"), (1, "dmma"), (2, "mixed")):
        t = C.c_double(0)
        eng.lib.kh_fp64_peak(mode, 20000, 148 * 8, C.c_void_p(scratch.data_ptr()), C.byref(t))
        print(f"peak {name}: {t.value:.2f} TFLOP/s", flush=True)

if __name__ == "__main__":
    if os.environ.get("KH_ZG_CHILD"):
        child(int(sys.argv[1]), int(sys.argv[2]))
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 98
        batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4140
        peaks()
        for v in os.environ.get("KH_ZG_VARIANTS", "9,1,0").split(","):
            env = dict(os.environ, KH_ZGEMM_VARIANT=v, KH_ZG_CHILD="1")
            subprocess.run([sys.executable, __file__, str(n), str(batch)], env=env)
