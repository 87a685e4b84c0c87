#!/usr/bin/env python
"""bench.py -- RCWA solves/sec (one solve = one (frequency, k-point), complex128) on B200.

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU implementation on the host cores
    python bench.py --workload bzi77-full [--gpus N]          # the literal configs[1] job (413 696 solves), strong scaling

One *step* = one pass of the hot path (Crystal.solve + poynting_flux_end) over one batch of synthetic sources per GPU.
Default workload = BASELINE.json configs[1]: the Brillouin-zone-integration grating stack (examples/bzi/bzi_animation.py:55-68),
7x7 harmonics (n = 98), 16 layers + 2 half spaces; a step covers `--kpoints` k-points of the 64x64 grid x 101 wavelengths per
GPU (weak scaling: every rank takes its own k-points, no data-path collective; NCCL only all-gathers the flux spectra).

Prints ONE JSON line: value = whole-job solves/s with inputs resident in HBM; e2e = the same through Crystal.solve_batch with
host buffers (H2D / D2H inside the timed region); roofline = the kernel that dominates the step BY TIME (achieved algorithmic
TFLOP/s from CUDA events around every launch, against the FP64 peak measured live), with the other kernels' fractions beside
it; cpu_baseline = the unmodified reference (baseline/_ref, `kind: reference`) or the oracle port on the box's host cores;
extra = device-timed solves/s for the 5x5 ... 15x15 bases and the drop-in scalar loop.
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

import workloads as wk  # noqa: E402  (synthetic geometry builders only: numpy, no oracle, no tests)

REF_DIR = os.path.join(ROOT, "baseline", "_ref")          # the unmodified reference, pip-installed (git-ignored, travels with gpurun)
WORKLOADS = {
    "bzi77": dict(desc="BASELINE configs[1]: BZI grating stack 7x7 harmonics, 64x64 k-grid x 101 wavelengths", pw=(7, 7), kpoints=41),
    "bzi77-full": dict(desc="BASELINE configs[1], the whole job: BZI grating stack 7x7 harmonics, all 64x64 k-points x 101 wavelengths = 413696 solves",
                       pw=(7, 7), kpoints=4096),
    "suh03": dict(desc="BASELINE configs[0]: README suh03 5x5 harmonics, [Scyl,S1,Scyl], 151 frequencies x kx sweep", pw=(5, 5), kpoints=64),
    "woodpile1111": dict(desc="BASELINE configs[2]: woodpile 11x11 harmonics, 200 k x 200 frequencies", pw=(11, 11), kpoints=8),
}


def make_workload(name, rank, step, kpoints):
    """Structure + the (wl, kp, pol) arrays of one step for one rank (deterministic)."""
    if name in ("bzi77", "bzi77-full"):
        st = wk.bzi_structure((7, 7))
        kg = wk.bzi_kgrid((64, 64)).reshape(2, -1)
        wls = 1 / np.linspace(0.8, 1.0, 101)
        first = 0 if name == "bzi77-full" else ((rank * 1009 + step) * kpoints) % kg.shape[1]
        ks = kg[:, (first + np.arange(kpoints)) % kg.shape[1]]
        wl = np.tile(wls, kpoints)
        kp = np.repeat(ks.T, len(wls), axis=0).astype(complex)
        pol = np.ones((wl.size, 2), dtype=complex)
    elif name == "suh03":
        st = wk.holey_pair(5, 128)
        freqs = np.linspace(0.49, 0.6, 151)
        kxs = np.linspace(0, 0.3 * np.pi, 256)
        first = ((rank * 101 + step) * kpoints) % 256
        kx = kxs[(first + np.arange(kpoints)) % 256]
        wl = np.tile(1 / freqs, kpoints)
        kp = np.stack([np.repeat(kx, 151), np.zeros(151 * kpoints)], 1).astype(complex)
        pol = np.tile(np.array([[1.0, 0.0]], dtype=complex), (wl.size, 1))
    elif name == "woodpile1111":
        st = wk.woodpile_structure((11, 11))
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


def nominal_flops_per_solve(st):
    """F = L_pat*209 n^3 + (Ls-1)*101.3 n^3 -- the NOMINAL dense count of the reference's literal schedule
    (every star product dense, eigensolver at 100 n^3).  The executed count is far lower (analytic uniform layers, collapsed
    runs, flux columns); it is reported separately from the kernels' own work counters."""
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


# --------------------------------------------------------------------------- CPU legs
# The reference's own recipe (examples/crystal_api/woodpile.py:12-22,132-136): BLAS pinned to one thread, KHEPRI_MT_ON=0,
# multiprocessing.Pool(ncores) over the sources.  kind "reference" = the UNMODIFIED reference from baseline/_ref driven
# through its own Crystal API (set_source; solve; poynting_flux_end); kind "port" = oracle/rcwa_oracle.py (numpy/LAPACK
# restatement, pinned against the reference's outputs) when baseline/_ref is absent.
def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "khepri"))


_ref_state = {}


def _cpu_worker_init(kind):
    for v in ("OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "OMP_NUM_THREADS"):
        os.environ[v] = "1"
    os.environ["KHEPRI_MT_ON"] = "0"
    try:
        from threadpoolctl import threadpool_limits
        global _limiter
        _limiter = threadpool_limits(1)
    except Exception:
        pass
    _ref_state["kind"] = kind
    if kind == "reference":
        sys.path.insert(0, REF_DIR)
        import khepri.crystal  # noqa: F401  (numba import: ~2 s, outside the timed region)


def _cpu_solve(args):
    st, wl, kp, te, tm = args
    if _ref_state.get("kind") == "reference":
        from khepri.crystal import Crystal as RefCrystal
        key = repr((st["pw"], st["stack"], sorted(st["layers"])))
        cl = _ref_state.get(key)
        if cl is None:
            cl = _ref_state[key] = wk.build_crystal(st, crystal_cls=RefCrystal)
        cl.set_source(wl, te, tm, kp=kp)
        cl.solve()
        R, T = cl.poynting_flux_end()
        return float(np.real(R)), float(np.real(T))
    from oracle import rcwa_oracle as orc
    return orc.solve_rt(st, wl, te, tm, kp=kp)


def cpu_time_sample(st, wl, kp, pol, nsolves, ncores, kind):
    """Returns (solves/s, seconds) for `nsolves` sources spread over the step's source list."""
    import multiprocessing as mp
    idx = np.linspace(0, wl.size - 1, nsolves).astype(int)
    jobs = [(st, float(wl[i]), (complex(kp[i, 0]), complex(kp[i, 1])), complex(pol[i, 0]), complex(pol[i, 1])) for i in idx]
    ctx = mp.get_context("fork")
    with ctx.Pool(ncores, initializer=_cpu_worker_init, initargs=(kind,)) as pool:
        pool.map(_cpu_solve, jobs[:ncores], chunksize=1)    # warm-up: imports, first touch, the reference's Crystal per worker
        t0 = time.perf_counter()
        pool.map(_cpu_solve, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return nsolves / dt, dt


def cpu_kind():
    return "reference" if reference_available() else "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    kind = cpu_kind()
    st, wl, kp, pol = make_workload(args.workload, 0, 0, args.kpoints)
    n = 2 * st["pw"][0] * st["pw"][1]
    per_step = max(ncores, int(args.cpu_solves or 8 * ncores))
    vals = []
    for step in range(args.warmup + args.steps):
        v, dt = cpu_time_sample(st, wl, kp, pol, per_step, ncores, kind)
        if step >= args.warmup:
            vals.append((v, dt))
    value = float(np.sum([per_step for _ in vals]) / np.sum([dt for _, dt in vals]))
    what = "unmodified reference (baseline/_ref) through khepri.crystal.Crystal" if kind == "reference" else "oracle port (baseline/_ref absent)"
    sample = f"{per_step} solves per step spread over the step's sources, Pool({ncores}) x 1 BLAS thread, {what}"
    line = {"impl": "reference", "metric": "RCWA solves/sec (freq x k-point, complex128)", "value": value, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([dt for _, dt in vals])),
            "higher_is_better": True, "scaling": "strong" if args.workload.endswith("-full") else "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "harmonics": list(st["pw"]), "n": n,
                       "solves_per_step_per_gpu": int(wl.size) if not args.workload.endswith("-full") else None, "solves_per_step": int(wl.size),
                       "kpoints_per_step_per_gpu": args.kpoints if not args.workload.endswith("-full") else None,
                       "method": "reference: eigen-decomposition per layer (numpy.linalg.eig), khepri.crystal.Crystal scalar loop",
                       "parallelism": f"Pool({ncores}) on the host cores, {per_step} solves sampled per step",
                       "l2": "n/a (CPU arm)", "results_finite": True},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": ncores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def profile_kernels(lib, fn):
    """Per-kernel CUDA-event times and algorithmic flops of everything `fn` launches (kh_profile_begin/end)."""
    lib.kh_profile_begin()
    fn()
    buf = C.create_string_buffer(1 << 16)
    lib.kh_profile_end(buf, len(buf))
    kernels = {}
    for ln in buf.value.decode().strip().splitlines():
        nm, cnt, ms, work = ln.split()
        kernels[nm] = dict(count=int(cnt), ms=float(ms), work=float(work))
    return kernels


def extras(eng, dev, fp64_peak, quick, rates_only=False):
    """Device-timed solves/s (inputs resident, flux only) for the 5x5 ... 15x15 bases of the north star, one short step each,
    and the drop-in scalar loop of README.md:55-59 (set_source; solve; poynting_flux_end per frequency, host API).
    rates_only (N > 1: every rank runs its own copy of each batch, no collective in here): the basis sweep without the
    scalar loop and the field maps."""
    import torch
    out = []

    def rate(name, st, wl, kp, pol, reps=1):
        cl = wk.build_crystal(st, eng)
        plan = cl._get_plan(False)
        w, k, p = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (wl, kp.astype(complex), pol.astype(complex)))
        res = eng.solve_batch(plan, w, k, p, want_flux=True, method=cl.method)         # warm-up (allocations, smem opt-in)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fb0 = eng.eig_fallbacks
        times = []
        for _ in range(max(3, reps)):              # median of at least three device-timed repetitions (host jitter on a busy multi-rank box)
            e0.record()
            res = eng.solve_batch(plan, w, k, p, want_flux=True, method=cl.method)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        fallbacks = (eng.eig_fallbacks - fb0) // len(times)          # sources re-solved with the eigen method (conditioning guard), inside the time
        rt = res["RT"].cpu().numpy()
        n = 2 * st["pw"][0] * st["pw"][1]
        v = len(wl) / (ms * 1e-3)
        out.append({"config": name, "harmonics": list(st["pw"]), "n": n, "solves": int(len(wl)), "ms": ms, "solves_per_s": v,
                    "nominal_tflops": nominal_flops_per_solve(st) * v / 1e12, "finite": bool(np.isfinite(rt).all()), "eig_fallbacks": int(fallbacks),
                    "max_abs_R_plus_T_minus_1": float(np.nanmax(np.abs(rt.sum(1) - 1)))})

    freqs = np.linspace(0.49, 0.6, 151)
    kx = np.linspace(0, 0.3 * np.pi, 64)
    wl = np.tile(1 / freqs, 64); kp = np.stack([np.repeat(kx, 151), np.zeros(151 * 64)], 1); pol = np.tile([[1.0, 0.0]], (wl.size, 1))
    rate("C1 suh03 5x5 (151 freqs x 64 kx)", wk.holey_pair(5, 128), wl, kp, pol, reps=2)
    rate("holey pair 9x9 (151 freqs x 8 kx)", wk.holey_pair(9, 128), wl[:151 * 8], kp[:151 * 8], pol[:151 * 8])
    fr = np.linspace(0.4 / 1.414, 0.65 / 1.414, 200); kxw = np.linspace(0, 0.99 * np.pi, 200)[:4]
    rate("C3 woodpile 11x11 (200 freqs x 4 kx)", wk.woodpile_structure((11, 11)), np.tile(1 / fr, 4), np.stack([np.repeat(kxw, 200), np.zeros(800)], 1), np.ones((800, 2)))
    if not quick:
        for pp, nf in ((13, 120), (15, 60)):
            fq = np.linspace(0.7, 0.83, nf)
            rate(f"direct {pp}x{pp} supercell basis, two pixmap layers ({nf} freqs)", wk.two_layer_structure(pp), 1 / fq, np.zeros((nf, 2)), np.tile([[1.0, 0.0]], (nf, 1)))
    if rates_only:
        return out
    # drop-in scalar loop (B = 1 per call, Stot materialised on the device, one D2H of (R, T) per frequency)
    st, srcs = wk.case_suh03()
    cl = wk.build_crystal(st, eng)
    for s in srcs[:3]:
        cl.set_source(**s); cl.solve(); cl.poynting_flux_end()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    acc = 0.0
    for s in srcs:
        cl.set_source(**s)
        cl.solve()
        acc += sum(cl.poynting_flux_end())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out.append({"config": "C1 suh03 5x5 drop-in SCALAR loop: 151 x (set_source; solve; poynting_flux_end), host API, wall clock",
                "harmonics": [5, 5], "n": 50, "solves": len(srcs), "ms": dt * 1e3, "solves_per_s": len(srcs) / dt, "finite": bool(np.isfinite(acc))})
    if not quick:
        # C5: field maps 9x9 on a 256 x 256 x 128 grid, 17 frequencies per batch (solve with retained eigenspaces + reconstruction;
        # the 13.7 GB of (E, H) stay on the device).  HBM-bound output: 6 nz nx ny 16 B per frequency.
        nf = 17
        st5, src5, _ = wk.case_fields(9, slices=4, res=128)
        xs = np.linspace(0, 1, 256); ys = np.linspace(0, 1, 256); zs = np.linspace(0.0001, 2.2, 128)
        X, Y = np.meshgrid(xs, ys, indexing="xy")
        cl5 = wk.build_crystal(st5, eng, fields=True)
        plan5 = cl5._get_plan(True)
        wl5 = 1 / np.linspace(0.49, 0.6, 51)[:nf]
        kp5 = np.zeros((nf, 2), dtype=complex); pol5 = np.tile([[1.0, 0.0]], (nf, 1)).astype(complex)
        inc = []
        for w in wl5:
            cl5.set_source(wavelength=float(w), te=1.0, tm=0.0)
            inc.append(np.hstack(cl5.get_source_as_field_vectors()))
        inc = np.asarray(inc)
        cl5.solve()
        F = None
        for rep in range(2):
            del F
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            solved = eng.solve_batch(plan5, wl5, kp5, pol5, want_flux=True, want_fields=True)
            e1.record()
            F = eng.fields(plan5, solved, wl5, kp5, inc, X.ravel(), Y.ravel(), zs, cl5.stack_positions, grid=(xs, ys))
            e2.record()
            torch.cuda.synchronize()
        ms_s, ms_f = e0.elapsed_time(e1), e1.elapsed_time(e2)
        gb = F.numel() * 16 / 1e9
        out.append({"config": f"C5 field maps 9x9, 256x256x128 grid, {nf} frequencies batched (device-resident output)", "harmonics": [9, 9], "n": 162,
                    "ms_solve_retained_eigenspaces": ms_s, "ms_field_kernels": ms_f, "ms_field_kernels_per_volume": ms_f / nf,
                    "volumes_per_s_incl_solve": nf / ((ms_s + ms_f) * 1e-3), "output_GB": gb, "field_kernels_GBps": gb / (ms_f * 1e-3),
                    "finite": bool(torch.isfinite(torch.view_as_real(F[0])).all().item())})
        del F
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from khepri_b200 import Engine
    from khepri_b200 import sharding

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
    full = args.workload.endswith("-full")          # strong scaling: the whole job is split over the ranks

    st, wl, kp, pol = make_workload(args.workload, rank, 0, args.kpoints)
    n = 2 * st["pw"][0] * st["pw"][1]
    cl = wk.build_crystal(st, eng)
    plan = cl._get_plan(False)
    nsteps = args.warmup + args.steps
    if full:
        lo, hi = sharding.shard_bounds(wl.size, world, rank)
        B_total, B = wl.size, hi - lo
        # warm-up steps of the standard step size (4141 solves of this rank's shard), then whole passes of the job
        nw = min(B, 4141)
        host_in = [(wl[lo:lo + nw], kp[lo:lo + nw], pol[lo:lo + nw])] * args.warmup + [(wl, kp, pol)] * args.steps
        dev_one = (torch.from_numpy(wl[lo:hi]).to(dev), torch.from_numpy(kp[lo:hi]).to(dev), torch.from_numpy(pol[lo:hi]).to(dev))
        dev_in = [tuple(t[:nw] for t in dev_one)] * args.warmup + [dev_one] * args.steps
    else:
        B = wl.size
        B_total = B * world

        def inputs(step):
            _, w, k, p = make_workload(args.workload, rank, step, args.kpoints)
            return (torch.from_numpy(w).to(dev), torch.from_numpy(k).to(dev), torch.from_numpy(p).to(dev)), (w, k, p)
        dev_in, host_in = zip(*[inputs(s) for s in range(nsteps)])
    def host_bounds(w, k):
        return float(np.min(w)), float(np.sqrt((np.abs(k) ** 2).sum(axis=1).max()))
    bounds = [host_bounds(h[0], h[1]) for h in host_in]        # computed with the inputs, outside the timed region
    method = eng._select_method(plan, None, None, False, cl.method, bounds[0])

    def step_resident(s):
        res = eng.solve_batch(plan, dev_in[s][0], dev_in[s][1], dev_in[s][2], want_flux=True, method=cl.method, bounds=bounds[s])
        if world > 1:
            return sharding.gather_spectra(res["RT"], B_total) if (full and s >= args.warmup) else (res["RT"] if full else _gather_equal(res["RT"]))
        return res["RT"]

    gather = [torch.empty((B, 2), dtype=torch.float64, device=dev) for _ in range(world)] if (world > 1 and not full) else None

    def _gather_equal(t):
        dist.all_gather(gather, t)
        return t

    def step_e2e(s):
        w, k, p = host_in[s]
        if full and s >= args.warmup:
            return sharding.sweep_sharded(cl, w, kps=k, te=p[:, 0], tm=p[:, 1])
        if full:
            return cl.solve_batch(w, kps=k, te=p[:, 0], tm=p[:, 1])
        R, T = cl.solve_batch(w, kps=k, te=p[:, 0], tm=p[:, 1])
        if world > 1:
            _gather_equal(torch.from_numpy(np.stack([R, T], 1)).to(dev))
        return R, T

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
        step_resident(s)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.kh_launch_count()
    fb_start = eng.eig_fallbacks
    ms_total = timed(step_resident, args.warmup, args.steps)
    fb_timed = eng.eig_fallbacks - fb_start
    launches = lib.kh_launch_count() - launches0
    # ---- e2e: the public API with host buffers (H2D of the sources, D2H of R and T, every step)
    step_e2e(0)
    ms_e2e = timed(step_e2e, args.warmup, args.steps)
    sampler.stop_flag = True
    sampler.join()
    # ---- the same steps again with CUDA events around every launch (which kernel bounds the step, and how close to its roofline)
    t_prof = [0.0]

    def prof_steps():
        t_prof[0] = timed(step_resident, args.warmup, args.steps)
    kernels = profile_kernels(lib, prof_steps)
    ms_prof = t_prof[0]
    tot_kernel_ms = sum(k["ms"] for k in kernels.values()) or 1.0
    rt = step_resident(nsteps - 1)
    rt = rt.cpu().numpy()
    finite = bool(np.isfinite(rt).all())
    # status words of one more step (0 = every solve clean: no zero pivot, no eigensolver failure, doubling bound respected)
    chk = eng.solve_batch(plan, dev_in[nsteps - 1][0], dev_in[nsteps - 1][1], dev_in[nsteps - 1][2], want_flux=True, method=cl.method, bounds=bounds[nsteps - 1])
    info_max = int(chk["info"].max().item())
    energy = float(np.abs(chk["RT"].cpu().numpy().sum(1) - 1).max())

    # N > 1: the 5x5 ... 15x15 basis sweep on every rank (no collective inside; one fixed-size all_gather of the device times after)
    sweep_ms, sweep_local = None, None
    if world > 1 and not args.no_extra and not full:
        K = 3 if args.quick_extra else 5
        try:
            sweep_local = extras(eng, dev, 0.0, args.quick_extra, rates_only=True)
        except Exception as exc:
            sweep_local = [{"config": "basis sweep", "error": repr(exc)}]
        v = torch.full((K,), float("nan"), dtype=torch.float64, device=dev)
        for i, e in enumerate(sweep_local[:K]):
            if "ms" in e and e.get("finite"):
                v[i] = e["ms"]
        allv = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        sweep_ms = torch.stack(allv).cpu()
        sweep_local = (sweep_local + [{"config": "(not run)"}] * K)[:K]

    if rank == 0:
        solves = B_total * args.steps
        value = solves / (ms_total * 1e-3)
        # FP64 peak measured live (MEASURED_PEAKS.json carries HBM and bf16 only): DFMA stream and DMMA stream
        scratch = torch.zeros(16, dtype=torch.float64, device=dev)
        peak = {}
        for mode, nm in ((0, "dfma"), (1, "dmma")):
            t = C.c_double()
            lib.kh_fp64_peak(mode, 20000, 148 * 8, C.c_void_p(scratch.data_ptr()), C.byref(t))
            peak[nm] = t.value
        fp64_peak = max(peak.values())
        table = {}
        for nm, k in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"]):
            e = {"share_of_kernel_time": round(k["ms"] / tot_kernel_ms, 4), "launches_per_step": k["count"] / args.steps, "avg_launch_ms": k["ms"] / k["count"]}
            if k["work"] > 0:
                e["achieved_tflops"] = k["work"] / (k["ms"] * 1e-3) / 1e12
                e["frac_of_fp64_peak"] = e["achieved_tflops"] / fp64_peak
            table[nm] = e
        dom = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
        roof = None
        if dom:
            d = kernels[dom]
            ach = d["work"] / (d["ms"] * 1e-3) / 1e12 if d["work"] > 0 else None
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "r02_dominant_traffic.json")
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                if tj.get("workload") == args.workload and tj.get("solves_per_step_per_gpu") == B and tj.get("kernel") == dom and tj.get("method") == method:
                    traffic = tj["dram_bytes_per_launch"]
            names = {"zgemm": "zgemm (batched complex128 DMMA GEMM, 8 M N K flop per product)", "zinv": "zinv (batched Gauss-Jordan inverse, 8 n^3)",
                     "zgeev_qr": "zgeev_qr (shifted QR iteration, nominal 50 n^3)", "zgeev_hess": "zgeev_hess (Hessenberg reduction, nominal 25 n^3)"}
            executed = sum(k["work"] for k in kernels.values())
            roof = {"bound": "tensor", "kernel": names.get(dom, dom), "dominant_by_time": dom, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (ach / fp64_peak) if ach else None, "traffic": traffic,
                    "peak_source": f"FP64 peak measured live by kh_fp64_peak (DFMA {peak['dfma']:.1f}, DMMA {peak['dmma']:.1f} TFLOP/s); MEASURED_PEAKS.json has no FP64 entry",
                    "launches": d["count"], "avg_launch_ms": d["ms"] / d["count"], "share_of_kernel_time": d["ms"] / tot_kernel_ms,
                    "profiled_ms_per_step": ms_prof / args.steps, "kernels": table,
                    "whole_step_executed_tflops": executed / (ms_prof * 1e-3) / 1e12,
                    "whole_step_executed_frac": executed / (ms_prof * 1e-3) / 1e12 / fp64_peak,
                    "executed_gflop_per_solve": executed / (B * args.steps) / 1e9,
                    "nominal_gflop_per_solve": nominal_flops_per_solve(st) / 1e9,
                    "nominal_note": "nominal = SURVEY 8(d) dense count of the reference's literal schedule; executed = sum of the kernels' own algorithmic counters"}
        cpu = None
        if world == 1 and not args.no_cpu:
            ncores = os.cpu_count() or 1
            kind = cpu_kind()
            per = max(ncores, int(args.cpu_solves or 8 * ncores))
            w0, k0, p0 = host_in[0]
            v, dt = cpu_time_sample(st, w0, k0, p0, per, ncores, kind)
            if not args.cpu_solves and dt < 8.0 and per < w0.size:          # bounded sample of about 12 s of CPU work
                per = min(w0.size, max(per, int(per * 12.0 / max(dt, 1e-3)) // ncores * ncores))
                v, dt = cpu_time_sample(st, w0, k0, p0, per, ncores, kind)
            what = "UNMODIFIED reference from baseline/_ref (khepri.crystal.Crystal: set_source; solve; poynting_flux_end)" if kind == "reference" \
                else "oracle = numpy/LAPACK restatement of the reference (baseline/_ref absent)"
            cpu = {"value": v, "unit": "solves/s", "cores": ncores, "kind": kind,
                   "sample": f"{per} solves spread over one step's sources in {dt:.1f} s, Pool({ncores}) x 1 BLAS thread, {what}"}
        extra = None
        if world == 1 and not args.no_extra and not full:
            try:
                extra = extras(eng, dev, fp64_peak, args.quick_extra)
            except Exception as exc:          # the headline line must survive a failing side measurement
                extra = [{"error": repr(exc)}]
        if sweep_ms is not None:              # N > 1: basis sweep, every rank its own copy of each batch; aggregate = N x solves / slowest rank
            extra = []
            for i, e in enumerate(sweep_local):
                ms = float(sweep_ms[:, i].max().item())
                if not np.isfinite(ms):
                    extra.append({"config": e.get("config"), "error": "a rank failed or skipped this basis"})
                    continue
                v = world * e["solves"] / (ms * 1e-3)
                e = dict(e, n_gpus=world, scaling="weak (every rank solves its own copy of the batch; max over ranks of the device time)",
                         ms=ms, solves=world * e["solves"], solves_per_s=v, nominal_tflops=e["nominal_tflops"] / e["solves_per_s"] * v)
                extra.append(e)
        line = {"metric": "RCWA solves/sec (freq x k-point, complex128)", "value": value, "unit": "solves/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong" if full else "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                "config": {"workload": WORKLOADS[args.workload]["desc"], "harmonics": list(st["pw"]), "n": n,
                           "solves_per_step_per_gpu": B, "solves_per_step": B_total, "kpoints_per_step_per_gpu": args.kpoints if not full else None,
                           "method": method + (" (slice power series + self star products; eigensolver only when eigenspaces are retained)" if method == "doubling" else ""),
                           "parallelism": f"dp{world} (independent (freq,k) solves sharded, NCCL all_gather of R,T only)",
                           "l2": "per-step working set (workspace of several GB) exceeds the 126 MB L2, no explicit flush",
                           "results_finite": finite, "results_info_max": info_max, "results_max_abs_R_plus_T_minus_1": energy,
                           "eig_fallbacks_in_timed_steps": int(fb_timed)},
                "clocks": sampler.summary(),
                "e2e": {"value": solves / (ms_e2e * 1e-3), "unit": "solves/s",
                        "h2d_bytes_per_step": int((B_total if full else B) * (8 + 32 + 32)) // (world if full else 1),
                        "d2h_bytes_per_step": int(B * 16)},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        if full:
            line["job_seconds"] = ms_total / args.steps * 1e-3
        if extra is not None:
            line["extra"] = extra
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
    ap.add_argument("--no-extra", action="store_true", help="skip the 5x5 ... 15x15 side measurements")
    ap.add_argument("--quick-extra", action="store_true", help="side measurements without the 13x13 / 15x15 bases")
    args = ap.parse_args()
    if args.kpoints <= 0 or args.workload.endswith("-full"):
        args.kpoints = WORKLOADS[args.workload]["kpoints"]
    if args.workload.endswith("-full") and args.impl == "cuda" and args.steps == 5:
        args.steps = 1                              # one pass of the whole job is the unit (warm-up steps are standard-size steps)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
