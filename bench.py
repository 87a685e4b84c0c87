#!/usr/bin/env python
"""bench.py -- RCWA solves/sec (one solve = one (frequency, k-point), complex128) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on the host cores

One *step* = one pass of the hot path (Crystal.solve + poynting_flux_end) over one batch of synthetic
sources per GPU.  Default workload = BASELINE.json configs[1]: the Brillouin-zone-integration grating
stack (examples/bzi/bzi_animation.py:55-68), 7x7 harmonics (n = 98), 16 layers + 2 half spaces; a step
covers `--kpoints` k-points of the 64x64 grid x 101 wavelengths per GPU (weak scaling: every rank takes
its own k-points, no data-path collective; NCCL only all-gathers the flux spectra).
Prints ONE JSON line (see the task contract): value = whole-job solves/s with inputs resident in HBM,
e2e = same through Crystal.solve_batch with host buffers, roofline for the dominant kernel (batched
DMMA GEMM) against the FP64 peak measured live, cpu_baseline = oracle port on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import cases  # noqa: E402  (synthetic geometry builders only; no oracle import here)

WORKLOADS = {
    "bzi77": dict(desc="BASELINE configs[1]: BZI grating stack 7x7 harmonics, 64x64 k-grid x 101 wavelengths", pw=(7, 7)),
    "suh03": dict(desc="BASELINE configs[0]: README suh03 5x5 harmonics, [Scyl,S1,Scyl], 151 frequencies x kx sweep", pw=(5, 5)),
    "woodpile1111": dict(desc="BASELINE configs[2]: woodpile 11x11 harmonics, 200 k x 200 frequencies", pw=(11, 11)),
}


def make_workload(name, rank, step, kpoints):
    """Structure + the (wl, kp, pol) arrays of one step for one rank (deterministic)."""
    if name == "bzi77":
        st = cases.bzi_structure((7, 7))
        kg = cases.bzi_kgrid((64, 64)).reshape(2, -1)
        wls = 1 / np.linspace(0.8, 1.0, 101)
        first = ((rank * 1009 + step) * kpoints) % kg.shape[1]
        ks = kg[:, (first + np.arange(kpoints)) % kg.shape[1]]
        wl = np.tile(wls, kpoints)
        kp = np.repeat(ks.T, len(wls), axis=0).astype(complex)
        pol = np.ones((wl.size, 2), dtype=complex)
    elif name == "suh03":
        st = cases.holey_pair(5, 128)
        freqs = np.linspace(0.49, 0.6, 151)
        kxs = np.linspace(0, 0.3 * np.pi, 256)
        first = ((rank * 101 + step) * kpoints) % 256
        kx = kxs[(first + np.arange(kpoints)) % 256]
        wl = np.tile(1 / freqs, kpoints)
        kp = np.stack([np.repeat(kx, 151), np.zeros(151 * kpoints)], 1).astype(complex)
        pol = np.tile(np.array([[1.0, 0.0]], dtype=complex), (wl.size, 1))
    elif name == "woodpile1111":
        st = cases.woodpile_structure((11, 11))
        freqs = np.linspace(0.4 / 1.414, 0.65 / 1.414, 200)
        kxs = np.linspace(0, 0.99 * np.pi, 200)[:100]
        first = ((rank * 37 + step) * kpoints) % 100
        kx = kxs[(first + np.arange(kpoints)) % 100]
        wl = np.tile(1 / freqs, kpoints)
        kp = np.stack([np.repeat(kx, 200), np.zeros(200 * kpoints)], 1).astype(complex)
        pol = np.ones((wl.size, 2), dtype=complex)
    else:
        raise SystemExit(f"unknown workload {name}")
    return st, wl, kp, pol


def algorithmic_flops_per_solve(st):
    """F = L_pat*209 n^3 + (Ls-1)*101.3 n^3 (nominal count of the reference's algorithm)."""
    n = 2 * st["pw"][0] * st["pw"][1]
    l_pat = sum(1 for name in set(st["stack"]) if st["layers"][name][0] == "pixmap")
    ls = len(st["stack"]) + 2
    return (l_pat * 209.0 + (ls - 1) * 101.3) * n ** 3


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------- CPU legs (oracle port)
def _cpu_worker_init():
    try:
        from threadpoolctl import threadpool_limits
        global _limiter
        _limiter = threadpool_limits(1)
    except Exception:
        pass


def _cpu_solve(args):
    from oracle import rcwa_oracle as orc
    st, wl, kp, te, tm = args
    return orc.solve_rt(st, wl, te, tm, kp=kp)


def cpu_time_sample(st, wl, kp, pol, nsolves, ncores):
    """The reference's own recipe (examples/crystal_api/woodpile.py:18-22,132-136): BLAS pinned to one
    thread, multiprocessing.Pool over the sources.  Returns (solves/s, seconds)."""
    import multiprocessing as mp
    idx = np.linspace(0, wl.size - 1, nsolves).astype(int)
    jobs = [(st, float(wl[i]), (complex(kp[i, 0]), complex(kp[i, 1])), complex(pol[i, 0]), complex(pol[i, 1])) for i in idx]
    ctx = mp.get_context("fork")
    with ctx.Pool(ncores, initializer=_cpu_worker_init) as pool:
        pool.map(_cpu_solve, jobs[:ncores])                 # warm-up: imports, first-touch
        t0 = time.perf_counter()
        pool.map(_cpu_solve, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return nsolves / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    st, wl, kp, pol = make_workload(args.workload, 0, 0, args.kpoints)
    n = 2 * st["pw"][0] * st["pw"][1]
    per_step = max(ncores, int(args.cpu_solves or 8 * ncores))
    vals = []
    for step in range(args.warmup + args.steps):
        v, dt = cpu_time_sample(st, wl, kp, pol, per_step, ncores)
        if step >= args.warmup:
            vals.append((v, dt))
    value = float(np.sum([per_step for _ in vals]) / np.sum([dt for _, dt in vals]))
    sample = f"{per_step} solves per step spread over the step's sources, Pool({ncores}) x 1 BLAS thread"
    line = {"impl": "reference", "metric": "RCWA solves/sec (freq x k-point, complex128)", "value": value, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([dt for _, dt in vals])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "harmonics": list(st["pw"]), "n": n},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": ncores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from khepri_b200 import Engine
    from tests.util import build_crystal

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = Engine(workspace_cap_bytes=int(args.workspace_gb * (1 << 30)))
    lib = eng.lib

    st, wl, kp, pol = make_workload(args.workload, rank, 0, args.kpoints)
    n = 2 * st["pw"][0] * st["pw"][1]
    B = wl.size
    cl = build_crystal(st, eng)
    plan = cl._get_plan(False)

    def device_inputs(step):
        _, w, k, p = make_workload(args.workload, rank, step, args.kpoints)
        return (torch.from_numpy(w).to(dev), torch.from_numpy(k).to(dev), torch.from_numpy(p).to(dev)), (w, k, p)

    nsteps = args.warmup + args.steps
    dev_in, host_in = zip(*[device_inputs(s) for s in range(nsteps)])
    gather = [torch.empty((B, 2), dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def step_resident(s):
        res = eng.solve_batch(plan, dev_in[s][0], dev_in[s][1], dev_in[s][2], want_flux=True)
        if world > 1:
            dist.all_gather(gather, res["RT"])
        return res["RT"]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, first, count):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(first, first + count):
            fn(s)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up, then the timed region (inputs resident in HBM)
    for s in range(args.warmup):
        last = step_resident(s)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.kh_launch_count()
    ms_total = timed(step_resident, args.warmup, args.steps)
    launches = lib.kh_launch_count() - launches0
    # ---- e2e: the public API with host buffers (H2D of the sources, D2H of R and T, every step)
    def step_e2e(s):
        w, k, p = host_in[s]
        R, T = cl.solve_batch(w, kps=k, te=p[:, 0], tm=p[:, 1])
        if world > 1:
            dist.all_gather(gather, torch.from_numpy(np.stack([R, T], 1)).to(dev))
        return R, T
    step_e2e(0)
    ms_e2e = timed(step_e2e, args.warmup, args.steps)
    sampler.stop_flag = True
    sampler.join()
    # ---- same steps again with per-kernel CUDA events (roofline of the dominant kernel)
    lib.kh_profile_begin()
    ms_prof = timed(step_resident, args.warmup, args.steps)
    buf = C.create_string_buffer(1 << 16)
    lib.kh_profile_end(buf, len(buf))
    kernels = {}
    for ln in buf.value.decode().strip().splitlines():
        nm, cnt, ms, work = ln.split()
        kernels[nm] = dict(count=int(cnt), ms=float(ms), work=float(work))
    tot_kernel_ms = sum(k["ms"] for k in kernels.values())
    # ---- sanity of the results of the last step (lossless stack: R + T = 1 is NOT expected for epse != 1 flux norm? it is: energy conservation)
    rt = step_resident(nsteps - 1).cpu().numpy()
    finite = bool(np.isfinite(rt).all())

    if rank == 0:
        solves = B * world * args.steps
        value = solves / (ms_total * 1e-3)
        # FP64 peak measured live (not in MEASURED_PEAKS.json): DFMA stream and DMMA stream
        scratch = torch.zeros(16, dtype=torch.float64, device=dev)
        peak = {}
        for mode, nm in ((0, "dfma"), (1, "dmma")):
            t = C.c_double()
            lib.kh_fp64_peak(mode, 20000, 148 * 8, C.c_void_p(scratch.data_ptr()), C.byref(t))
            peak[nm] = t.value
        fp64_peak = max(peak.values())
        dom = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
        gem = kernels.get("zgemm")
        roof = None
        if gem:
            ach = gem["work"] / (gem["ms"] * 1e-3) / 1e12
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "r01_zgemm_traffic.json")
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                if tj.get("workload") == args.workload and tj.get("solves_per_step_per_gpu") == B:
                    traffic = tj["dram_bytes_per_launch"]
            roof = {"bound": "tensor", "kernel": "zgemm (batched complex128 DMMA GEMM)", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach / fp64_peak, "traffic": traffic,
                    "peak_source": f"FP64 peak measured live by kh_fp64_peak (DFMA {peak['dfma']:.1f}, DMMA {peak['dmma']:.1f} TFLOP/s); MEASURED_PEAKS.json has no FP64 entry",
                    "launches": gem["count"], "avg_launch_ms": gem["ms"] / gem["count"], "share_of_kernel_time": gem["ms"] / tot_kernel_ms,
                    "dominant_by_time": dom, "profiled_ms_per_step": ms_prof / args.steps,
                    "kernel_time_shares": {k: round(v["ms"] / tot_kernel_ms, 4) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])},
                    "whole_solve_algorithmic_tflops": algorithmic_flops_per_solve(st) * value / 1e12}
        cpu = None
        if world == 1 and not args.no_cpu:
            ncores = os.cpu_count() or 1
            per = max(ncores, int(args.cpu_solves or 16 * ncores))
            v, dt = cpu_time_sample(st, host_in[0][0], host_in[0][1], host_in[0][2], per, ncores)
            if not args.cpu_solves and dt < 8.0 and per < B:          # bounded sample of about 12 s of CPU work
                per = min(B, max(per, int(per * 12.0 / max(dt, 1e-3)) // ncores * ncores))
                v, dt = cpu_time_sample(st, host_in[0][0], host_in[0][1], host_in[0][2], per, ncores)
            cpu = {"value": v, "unit": "solves/s", "cores": ncores, "kind": "port",
                   "sample": f"{per} solves spread over one step's sources in {dt:.1f} s, Pool({ncores}) x 1 BLAS thread (oracle = numpy/LAPACK restatement of the reference)"}
        line = {"metric": "RCWA solves/sec (freq x k-point, complex128)", "value": value, "unit": "solves/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                "config": {"workload": WORKLOADS[args.workload]["desc"], "harmonics": list(st["pw"]), "n": n,
                           "solves_per_step_per_gpu": B, "kpoints_per_step_per_gpu": args.kpoints,
                           "parallelism": f"dp{world} (independent (freq,k) solves sharded, NCCL all_gather of R,T only)",
                           "l2": "per-step working set (workspace of several GB) exceeds the 126 MB L2, no explicit flush",
                           "results_finite": finite},
                "clocks": sampler.summary(),
                "e2e": {"value": solves / (ms_e2e * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": int(B * (8 + 32 + 32)), "d2h_bytes_per_step": int(B * 16)},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="bzi77", choices=sorted(WORKLOADS))
    ap.add_argument("--kpoints", type=int, default=0, help="k-points per step per GPU (x wavelengths = solves per step); default 41 (bzi77), 64 (suh03), 8 (woodpile1111)")
    ap.add_argument("--workspace-gb", type=float, default=80.0)
    ap.add_argument("--cpu-solves", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.kpoints <= 0:
        args.kpoints = {"bzi77": 41, "suh03": 64, "woodpile1111": 8}[args.workload]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
