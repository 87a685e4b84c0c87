"""Shared helpers for the parity tests."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EMU_LIB = os.path.join(ROOT, "tests", "hostemu", "libkh_hostemu.so")

# Every parity test runs against two builds of the SAME kernel sources through the SAME C ABI:
#   "cuda": the product library on a B200 (marked gpu)
#   "emu":  tests/hostemu -- the sources compiled as host C++ with one virtual thread per CTA.  Test
#           infrastructure only: it checks the kernel logic on machines without a GPU.
BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]

_engines = {}


def engine(kind):
    if kind in _engines:
        return _engines[kind]
    from khepri_b200 import Engine
    if kind == "cuda":
        import torch
        assert torch.cuda.is_available(), "gpu test without a CUDA device"
        eng = Engine()
    else:
        srcs = [os.path.join(ROOT, "khepri_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "khepri_b200", "csrc"))]
        if not os.path.exists(EMU_LIB) or any(os.path.getmtime(s) > os.path.getmtime(EMU_LIB) for s in srcs):
            subprocess.run(["sh", os.path.join(ROOT, "tests", "hostemu", "build.sh")], check=True)
        eng = Engine(device="cpu", lib_path=EMU_LIB)
    _engines[kind] = eng
    return eng


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


# both ways a patterned layer gets its S-matrix (Engine._select_method): the eigen-decomposition, and "auto" = the doubling method
# with its conditioning guard (sources flagged by the device are solved again with the eigen method)
METHODS = ["eig", "auto"]


def build_crystal(st, eng, fields=False, method="auto"):
    import workloads
    return workloads.build_crystal(st, eng, fields=fields, method=method)


def sweep_sources(cl, srcs, **kw):
    """Batched sweep from a list of set_source kwargs."""
    wl = [s["wavelength"] for s in srcs]
    te = [s.get("te", 1.0) for s in srcs]
    tm = [s.get("tm", 1.0) for s in srcs]
    if srcs[0].get("kp") is not None:
        return cl.solve_batch(wl, kps=[s["kp"] for s in srcs], te=te, tm=tm, **kw)
    return cl.solve_batch(wl, te=te, tm=tm, theta=[s.get("theta", 0.0) for s in srcs], phi=[s.get("phi", 0.0) for s in srcs], **kw)
