"""Host-side plumbing between torch tensors and the C ABI (include/khepri_b200.h).

PyTorch is used only for device memory, streams and (in bench.py) torch.distributed; all compute
happens in the hand-written sm_100a kernels of ``khepri_b200/lib/libkhepri_b200.so``.

There is no CPU fallback.  ``Engine()`` raises when the CUDA library or a CUDA device is missing.
(The ``device="cpu"`` + ``lib_path=`` combination exists solely so that tests/hostemu can drive a
host-emulation build of the very same kernel sources through this code; the package never selects
it by itself.)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import KhepriError, LayerDesc, Outputs, check

_c128 = torch.complex128
_f64 = torch.float64


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class Plan:
    """Device-side description of a Crystal's geometry (crystal.py:42-164)."""

    def __init__(self, engine, pw, g, epsi, epse, layers, stack, Nb=0, glhs=None, grhs=None):
        self.engine = engine
        self.pw = (int(pw[0]), int(pw[1]))
        self.N = self.pw[0] * self.pw[1]
        self.n = 2 * self.N
        self.layers = layers
        self.stack = [int(s) for s in stack]
        self.Ls = len(self.stack)
        self.nL = len(layers)
        dev = engine.device
        g_host = np.ascontiguousarray(g, dtype=np.float64)
        # ky depends on the q index only (b1 has no y component: square / rectangular lattices) -> separable field maps
        gy = g_host[1].reshape(self.pw[1], self.pw[0]) if g_host.shape == (2, self.N) else None
        self.grid_separable = bool(gy is not None and self.pw[1] <= 16 and np.all(gy == gy[:, :1]))
        self.g = torch.as_tensor(g_host, device=dev)
        assert self.g.shape == (2, self.N), "g-vectors must have shape (2, N)"
        self.glhs = torch.as_tensor(np.ascontiguousarray(glhs, dtype=np.float64), device=dev) if glhs is not None else None
        self.grhs = torch.as_tensor(np.ascontiguousarray(grhs, dtype=np.float64), device=dev) if grhs is not None else None
        self._keep = []
        self.gmax = float(np.sqrt((g_host ** 2).sum(axis=0)).max())
        self.eps_bound = max([float(L.get("eps_bound", 0.0)) for L in layers] + [0.0])
        self.has_patterned = any(int(L["kind"]) == _lib.LAYER_PIXMAP for L in layers)
        # the spectral bound of the doubling method holds for dielectrics; layers without the information count as unsafe
        self.spectrum_bounded = all(float(L.get("eps_min_real", -1.0)) >= 0.5 for L in layers if int(L["kind"]) == _lib.LAYER_PIXMAP)
        # deepest patterned slab the flux-only solve sees: runs of one repeated layer are merged into one layer of m x depth (plan_view)
        self.max_patterned_depth, run, prev = 0.0, 0.0, None
        for li in self.stack:
            L = layers[li]
            if int(L["kind"]) != _lib.LAYER_PIXMAP:
                run, prev = 0.0, None
                continue
            run = run + float(L.get("depth", 0.0)) if li == prev else float(L.get("depth", 0.0))
            prev = li
            self.max_patterned_depth = max(self.max_patterned_depth, run)
        self._method_state = None
        descs = (LayerDesc * len(layers))()
        for i, L in enumerate(layers):
            d = descs[i]
            d.kind = int(L["kind"])
            eps = complex(L.get("eps", 1.0))
            d.eps_re, d.eps_im = eps.real, eps.imag
            d.depth = float(L.get("depth", 0.0))
            d.retain = int(bool(L.get("retain", False)))
            d.ext_base = int(L.get("ext_base", -1))
            d.ext_mode = int(L.get("ext_mode", 0))
            d.C_dev = d.IC_dev = None
            if d.kind == _lib.LAYER_PIXMAP:
                Cm = L["C"].to(device=dev, dtype=_c128).contiguous()
                ICm = L["IC"].to(device=dev, dtype=_c128).contiguous()
                self._keep += [Cm, ICm]
                d.C_dev, d.IC_dev = Cm.data_ptr(), ICm.data_ptr()
        stack_arr = (C.c_int * self.Ls)(*self.stack)
        handle = C.c_void_p()
        lib = engine.lib
        check(lib, lib.kh_plan_create(C.byref(handle), self.pw[0], self.pw[1], _ptr(self.g),
                                      complex(epsi).real, complex(epsi).imag, complex(epse).real, complex(epse).imag,
                                      len(layers), descs, self.Ls, stack_arr, int(Nb), _ptr(self.glhs), _ptr(self.grhs)),
              "kh_plan_create")
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.engine.lib.kh_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Engine:
    _default = None

    def __init__(self, device=None, lib_path=None, workspace_cap_bytes=None):
        if device is None:
            if not torch.cuda.is_available():
                raise KhepriError("khepri_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda" and lib_path is None:
            raise KhepriError("khepri_b200 only runs on CUDA devices; there is no CPU fallback.")
        self.lib = _lib.bind(lib_path)
        self._ws = None
        self.eig_fallbacks = 0          # sources that method "auto" solved again with the eigen method (ill-conditioned doubling)
        self.workspace_cap_bytes = workspace_cap_bytes
        self.launch_count = 0
        self.doubling_theta = 10.0      # largest |lambda k0 d| of one slice of the doubling method (see _select_method)

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = cls()
        return cls._default

    # ------------------------------------------------------------------ helpers
    def stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(None)

    def _cap(self):
        if self.workspace_cap_bytes is not None:
            return int(self.workspace_cap_bytes)
        if self.device.type == "cuda":
            free, _total = torch.cuda.mem_get_info(self.device)
            held = self._ws.numel() if self._ws is not None else 0
            return int(min(96 << 30, 0.6 * (free + held)))        # of the B200's 180 GB: large chunks keep the latency-bound kernels in full waves
        return 1 << 30

    def workspace(self, nbytes):
        nbytes = int(nbytes)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def to_dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        arr = np.ascontiguousarray(a)
        t = torch.from_numpy(arr)
        if self.device.type == "cuda":
            # large inputs go through pinned memory (asynchronous DMA); for a few KB the page-locking costs more than the copy
            t = t.pin_memory().to(self.device, non_blocking=True) if arr.nbytes >= (1 << 16) else t.to(self.device)
        return t.to(dtype).contiguous()

    # ------------------------------------------------------------------ convolution matrix
    def convmat(self, pixmaps, pw, return_coefficients=False):
        """tools.convolution_matrix for a stack of pixmaps [L, Nx, Ny] (real or complex)."""
        pix = pixmaps if isinstance(pixmaps, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pixmaps))
        if pix.dim() == 2:
            pix = pix[None]
        is_complex = pix.is_complex()
        pix = pix.to(device=self.device, dtype=_c128 if is_complex else _f64).contiguous()
        L, Nx, Ny = pix.shape
        P, Q = int(pw[0]), int(pw[1])
        if (P > 1 and Nx // 2 + P - 1 >= Nx) or (Q > 1 and Ny // 2 + Q - 1 >= Ny):
            raise IndexError("harmonic differences exceed the Fourier grid")      # as the reference's gather would
        N = P * Q
        Cm = torch.empty((L, N, N), dtype=_c128, device=self.device)
        F = torch.empty((L, 2 * P - 1, 2 * Q - 1), dtype=_c128, device=self.device)
        wb = self.lib.kh_convmat_work_bytes(L, Nx, Ny, P, Q)
        ws = self.workspace(wb)
        check(self.lib, self.lib.kh_convmat(L, Nx, Ny, int(is_complex), _ptr(pix), P, Q, _ptr(Cm), _ptr(F), _ptr(ws), ws.numel(), self.stream()), "kh_convmat")
        return (Cm, F) if return_coefficients else Cm

    def toeplitz_gather(self, F, pw):
        """tools.convolution_matrix_fourier (bit-exact index gather)."""
        Ft = self.to_dev(F, _c128)
        Nx, Ny = Ft.shape
        P, Q = int(pw[0]), int(pw[1])
        Cm = torch.empty((P * Q, P * Q), dtype=_c128, device=self.device)
        err = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(self.lib, self.lib.kh_toeplitz_gather(_ptr(Ft), Nx, Ny, P, Q, _ptr(Cm), _ptr(err), self.stream()), "kh_toeplitz_gather")
        if int(err.item()):
            raise IndexError("harmonic differences exceed the Fourier grid")
        return Cm

    # ------------------------------------------------------------------ dense primitives
    def zgemm(self, A, B, transA=False, alpha=1.0):
        A = self.to_dev(A, _c128)
        B = self.to_dev(B, _c128)
        if A.dim() == 2:
            A = A[None]
        if B.dim() == 2:
            B = B[None]
        batch = max(A.shape[0], B.shape[0])
        M, K = (A.shape[2], A.shape[1]) if transA else (A.shape[1], A.shape[2])
        N = B.shape[2]
        assert B.shape[1] == K
        out = torch.empty((batch, M, N), dtype=_c128, device=self.device)
        sA = 0 if A.shape[0] == 1 and batch > 1 else A.shape[1] * A.shape[2]
        sB = 0 if B.shape[0] == 1 and batch > 1 else B.shape[1] * B.shape[2]
        check(self.lib, self.lib.kh_zgemm_batched(batch, M, N, K, int(transA), _ptr(A), A.shape[2], sA, _ptr(B), B.shape[2], sB,
                                                  _ptr(out), N, M * N, float(alpha), self.stream()), "kh_zgemm_batched")
        return out

    def zinv(self, A, return_info=False):
        A = self.to_dev(A, _c128)
        squeeze = A.dim() == 2
        if squeeze:
            A = A[None]
        batch, n, _ = A.shape
        out = torch.empty_like(A)
        info = torch.zeros(batch, dtype=torch.int32, device=self.device)
        wb = self.lib.kh_zinv_work_bytes(batch, n)
        ws = torch.empty(max(wb, 16), dtype=torch.uint8, device=self.device)
        check(self.lib, self.lib.kh_zinv_batched(batch, n, _ptr(A), _ptr(out), _ptr(info), _ptr(ws), wb, self.stream()), "kh_zinv_batched")
        out = out[0] if squeeze else out
        return (out, info) if return_info else out

    def zgeev(self, A):
        A = self.to_dev(A, _c128)
        if A.dim() == 2:
            A = A[None]
        batch, n, _ = A.shape
        w = torch.empty((batch, n), dtype=_c128, device=self.device)
        W = torch.empty((batch, n, n), dtype=_c128, device=self.device)
        info = torch.zeros(batch, dtype=torch.int32, device=self.device)
        wb = self.lib.kh_zgeev_work_bytes(batch, n)
        ws = self.workspace(wb)
        check(self.lib, self.lib.kh_zgeev_batched(batch, n, _ptr(A), _ptr(w), _ptr(W), _ptr(ws), ws.numel(), _ptr(info), self.stream()), "kh_zgeev_batched")
        return w, W, info

    # ------------------------------------------------------------------ the batched solve
    # ------------------------------------------------------------------ patterned-layer method
    def _select_method(self, plan, wl, kp, want_fields, method, bounds=None):
        """Picks how patterned layers get their S-matrix (kh_plan_set_method).  "eig": eigen-decomposition (the only
        choice when eigenspaces are retained); "doubling": slice series + self star products, GEMMs only.  "auto" takes
        doubling whenever no eigenspace is retained.  For doubling the host supplies kappa >= k0 sqrt(rho(Omega^2)):
        rho is bounded from the largest in-plane wavevector of the basis and the largest permittivity (the device checks
        the bound per solve and raises info bit 2)."""
        import os
        if method in (None, "auto"):
            method = os.environ.get("KHEPRI_B200_METHOD", "auto")          # developer switch for A/B runs of bench.py
        if method == "auto":
            method = "doubling" if (plan.has_patterned and not want_fields) else "eig"
        if method not in ("eig", "doubling"):
            raise ValueError(f"unknown method {method!r}")
        if method == "doubling" and (want_fields or not plan.has_patterned):
            method = "eig"
        if method == "eig":
            state = ("eig",)
            if plan._method_state != state:
                check(self.lib, self.lib.kh_plan_set_method(plan.handle, _lib.METHOD_EIG, 0.0, 0.0), "kh_plan_set_method")
                plan._method_state = state
            return method
        if bounds is not None:                  # (min wavelength, max |kp|) supplied by the caller: no device round trip
            wl_min, kp_max = float(bounds[0]), float(bounds[1])
        elif isinstance(wl, torch.Tensor):
            wl_min = float(wl.min().item())
            kp_max = float(kp.abs().pow(2).sum(dim=-1).max().sqrt().item())
        else:
            wl_min = float(np.min(wl))
            kpa = np.abs(np.asarray(kp)).reshape(-1, 2)
            kp_max = float(np.sqrt((kpa ** 2).sum(axis=1).max()))
        k0 = 2.0 * np.pi / wl_min
        kappa = float(np.sqrt(1.25 * (kp_max + plan.gmax) ** 2 + plan.eps_bound * k0 * k0))
        theta = float(os.environ.get("KHEPRI_B200_THETA", self.doubling_theta))
        state = ("doubling", kappa, theta)
        if plan._method_state != state:
            check(self.lib, self.lib.kh_plan_set_method(plan.handle, _lib.METHOD_DOUBLING, kappa, theta), "kh_plan_set_method")
            plan._method_state = state
        return method

    def solve_batch(self, plan, wl, kp, pol=None, want_S=False, want_flux=True, want_orders=False, want_fields=False, chunk=None, method=None, bounds=None):
        """Crystal.solve (+ poynting_flux_end) for B sources.  wl [B], kp [B,2] complex, pol [B,2] = (te, tm).

        Returns a dict of DEVICE tensors: RT [B,2], orders [B,2,N], Stot [B,2,2,n,n], info [B] and, with
        want_fields, prefix/suffix [B,Ls,2,2,n,n], W/V [B,nL,n,n], L [B,nL,n].
        method: "auto" | "eig" | "doubling" (see _select_method).  bounds = (min wavelength, max |kp|) of the batch, optional:
        with DEVICE inputs it saves the reduction + host read the doubling method otherwise needs to size its series.
        """
        used = None
        if (wl.numel() if isinstance(wl, torch.Tensor) else np.size(wl)) > 0:
            used = self._select_method(plan, wl, kp, want_fields, method, bounds)
        wl_d = self.to_dev(np.asarray(wl, dtype=np.float64).reshape(-1) if not isinstance(wl, torch.Tensor) else wl.reshape(-1), _f64)
        B = wl_d.numel()
        kp_d = self.to_dev(np.asarray(kp, dtype=np.complex128).reshape(B, 2) if not isinstance(kp, torch.Tensor) else kp.reshape(B, 2), _c128)
        pol_d = None
        if want_flux:
            assert pol is not None, "flux needs the (te, tm) polarisation"
            pol_d = self.to_dev(np.asarray(pol, dtype=np.complex128).reshape(B, 2) if not isinstance(pol, torch.Tensor) else pol.reshape(B, 2), _c128)
        n, N, Ls, nL = plan.n, plan.N, plan.Ls, plan.nL
        dev = self.device
        res = {"info": torch.zeros(B, dtype=torch.int32, device=dev)}
        out = Outputs()
        out.info_dev = res["info"].data_ptr()
        flags = 0
        if want_S:
            res["Stot"] = torch.empty((B, 2, 2, n, n), dtype=_c128, device=dev)
            out.Stot_dev = res["Stot"].data_ptr()
            flags |= _lib.WANT_STOT
        if want_flux:
            res["RT"] = torch.empty((B, 2), dtype=_f64, device=dev)
            out.RT_dev = res["RT"].data_ptr()
            flags |= _lib.WANT_FLUX
            if want_orders:
                res["orders"] = torch.empty((B, 2, N), dtype=_f64, device=dev)
                out.orders_dev = res["orders"].data_ptr()
        if want_fields:
            res["prefix"] = torch.empty((B, Ls, 2, 2, n, n), dtype=_c128, device=dev)
            res["suffix"] = torch.empty((B, Ls, 2, 2, n, n), dtype=_c128, device=dev)
            res["W"] = torch.zeros((B, nL, n, n), dtype=_c128, device=dev)
            res["V"] = torch.zeros((B, nL, n, n), dtype=_c128, device=dev)
            res["L"] = torch.zeros((B, nL, n), dtype=_c128, device=dev)
            out.prefix_dev, out.suffix_dev = res["prefix"].data_ptr(), res["suffix"].data_ptr()
            out.W_dev, out.V_dev, out.L_dev = res["W"].data_ptr(), res["V"].data_ptr(), res["L"].data_ptr()
            flags |= _lib.WANT_FIELDS
        if B == 0:
            return res
        lib = self.lib
        want = int(chunk) if chunk else B
        need = lib.kh_solve_workspace_bytes(plan.handle, want, flags)
        if self._ws is not None and self._ws.numel() >= need:
            ws = self._ws                          # (no cudaMemGetInfo on the hot path: 1.5 ms per call, most of a scalar solve)
        else:
            cap = self._cap()
            one = lib.kh_solve_workspace_bytes(plan.handle, 1, flags)
            if one > cap:
                raise KhepriError(f"one solve needs {one} bytes of workspace, cap is {cap}")
            ws = self.workspace(min(need, cap) if not chunk else need)
        ws_bytes = min(ws.numel(), need) if chunk else ws.numel()
        check(lib, lib.kh_solve_batch(plan.handle, B, _ptr(wl_d), _ptr(kp_d), _ptr(pol_d), C.byref(out), _ptr(ws), ws_bytes, self.stream()),
              "kh_solve_batch")
        # "auto": sources the doubling method reports as unsafe -- a self star product through a sub-slab resonance (info bit 3) or a
        # spectrum beyond the host's bound (bit 2: metallic gratings, whose inverse convolution matrix is not bounded by max |eps|) --
        # are solved again with the eigen method.  Costs one host read of the status words, which batches that can raise neither bit
        # (no layer deep enough for a self star product, dielectrics only) skip.
        if used == "doubling" and method in (None, "auto") and \
                (not plan.spectrum_bounded or plan._method_state[1] * plan.max_patterned_depth > 2.0 * plan._method_state[2]):
            flagged = (res["info"] & 12) != 0
            if bool(flagged.any().item()):
                idx = flagged.nonzero().flatten()
                sub = self.solve_batch(plan, wl_d[idx], kp_d[idx], pol_d[idx] if pol_d is not None else None, want_S=want_S, want_flux=want_flux,
                                       want_orders=want_orders, want_fields=want_fields, chunk=chunk, method="eig")
                for k, v in sub.items():
                    res[k][idx] = v
                self.eig_fallbacks += int(idx.numel())
        return res

    def star(self, SA, SB):
        """alternative.redheffer_product on stacks [B,2,2,n,n]."""
        SA = self.to_dev(SA, _c128)
        SB = self.to_dev(SB, _c128)
        squeeze = SA.dim() == 4
        if squeeze:
            SA, SB = SA[None], SB[None]
        B, n = SA.shape[0], SA.shape[-1]
        SO = torch.empty_like(SA)
        wb = self.lib.kh_star_workspace_bytes(B, n)
        ws = self.workspace(wb)
        check(self.lib, self.lib.kh_star_batch(B, n, _ptr(SA), _ptr(SB), _ptr(SO), _ptr(ws), ws.numel(), self.stream()), "kh_star_batch")
        return SO[0] if squeeze else SO

    def flux(self, plan, Stot, wl, kp, pol, want_orders=False):
        Stot = self.to_dev(Stot, _c128)
        if Stot.dim() == 4:
            Stot = Stot[None]
        B = Stot.shape[0]
        wl_d = self.to_dev(np.asarray(wl, dtype=np.float64).reshape(B), _f64)
        kp_d = self.to_dev(np.asarray(kp, dtype=np.complex128).reshape(B, 2), _c128)
        pol_d = self.to_dev(np.asarray(pol, dtype=np.complex128).reshape(B, 2), _c128)
        RT = torch.empty((B, 2), dtype=_f64, device=self.device)
        orders = torch.empty((B, 2, plan.N), dtype=_f64, device=self.device) if want_orders else None
        check(self.lib, self.lib.kh_flux_batch(plan.handle, B, _ptr(Stot), _ptr(wl_d), _ptr(kp_d), _ptr(pol_d), _ptr(RT), _ptr(orders), self.stream()),
              "kh_flux_batch")
        return (RT, orders) if want_orders else RT

    def fields(self, plan, solved, wl, kp, inc, x, y, z, stack_positions, grid=None):
        """(Ex,Ey,Ez,Hx,Hy,Hz) at points (x[p], y[p]) and depths z for every solve of a want_fields
        solve -> DEVICE tensor [B, nz, 6, npts].  inc [B, 2, n] = incident (E, H) Fourier vectors.
        grid=(xs, ys): the points are the rectangular grid x = xs[ix], y = ys[iy] (p = iy * nx + ix) and the
        separable transform is used when the lattice allows it (plan.grid_separable)."""
        B = solved["prefix"].shape[0]
        if grid is not None and plan.grid_separable:
            return self._fields_grid(plan, solved, wl, kp, inc, grid[0], grid[1], z, stack_positions)
        wl_d = self.to_dev(np.asarray(wl, dtype=np.float64).reshape(B), _f64)
        kp_d = self.to_dev(np.asarray(kp, dtype=np.complex128).reshape(B, 2), _c128)
        inc_d = self.to_dev(np.asarray(inc, dtype=np.complex128).reshape(B, 2, plan.n), _c128)
        x_d = self.to_dev(np.asarray(x, dtype=np.float64).reshape(-1), _f64)
        y_d = self.to_dev(np.asarray(y, dtype=np.float64).reshape(-1), _f64)
        z_h = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(-1))
        zp_h = np.ascontiguousarray(np.asarray(stack_positions, dtype=np.float64).reshape(-1))
        assert zp_h.size == plan.Ls + 1, "stack_positions must hold Ls+1 interface positions"
        npts, nz = x_d.numel(), z_h.size
        assert y_d.numel() == npts
        F = torch.empty((B, nz, 6, npts), dtype=_c128, device=self.device)
        out = Outputs()
        out.prefix_dev, out.suffix_dev = solved["prefix"].data_ptr(), solved["suffix"].data_ptr()
        out.W_dev, out.V_dev, out.L_dev = solved["W"].data_ptr(), solved["V"].data_ptr(), solved["L"].data_ptr()
        wb = self.lib.kh_fields_workspace_bytes(plan.handle, B, npts, nz)
        ws = self.workspace(wb)
        dp = C.POINTER(C.c_double)
        check(self.lib, self.lib.kh_fields_batch(plan.handle, B, _ptr(wl_d), _ptr(kp_d), _ptr(inc_d), C.byref(out), _ptr(x_d), _ptr(y_d), npts,
                                                 z_h.ctypes.data_as(dp), nz, zp_h.ctypes.data_as(dp), _ptr(F), _ptr(ws), ws.numel(), self.stream()),
              "kh_fields_batch")
        return F

    def fields_fourier(self, plan, solved, wl, kp, inc, z, stack_positions):
        """Fourier fields (sx, sy, sz, ux, uy, uz) at depths z for every solve of a want_fields solve
        (crystal.py:234-277) -> DEVICE tensor [B, nz, 6, N]."""
        B = solved["prefix"].shape[0]
        wl_d = self.to_dev(np.asarray(wl, dtype=np.float64).reshape(B), _f64)
        kp_d = self.to_dev(np.asarray(kp, dtype=np.complex128).reshape(B, 2), _c128)
        inc_d = self.to_dev(np.asarray(inc, dtype=np.complex128).reshape(B, 2, plan.n), _c128)
        z_h = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(-1))
        zp_h = np.ascontiguousarray(np.asarray(stack_positions, dtype=np.float64).reshape(-1))
        assert zp_h.size == plan.Ls + 1, "stack_positions must hold Ls+1 interface positions"
        nz = z_h.size
        S = torch.empty((B, nz, 6, plan.n // 2), dtype=_c128, device=self.device)
        out = Outputs()
        out.prefix_dev, out.suffix_dev = solved["prefix"].data_ptr(), solved["suffix"].data_ptr()
        out.W_dev, out.V_dev, out.L_dev = solved["W"].data_ptr(), solved["V"].data_ptr(), solved["L"].data_ptr()
        wb = self.lib.kh_fields_workspace_bytes(plan.handle, B, 1, nz)
        ws = self.workspace(wb)
        dp = C.POINTER(C.c_double)
        check(self.lib, self.lib.kh_fields_fourier_batch(plan.handle, B, _ptr(wl_d), _ptr(kp_d), _ptr(inc_d), C.byref(out),
                                                         z_h.ctypes.data_as(dp), nz, zp_h.ctypes.data_as(dp), _ptr(S), _ptr(ws), ws.numel(), self.stream()),
              "kh_fields_fourier_batch")
        return S

    def idft(self, s, kx, ky, x, y):
        """fourier.idft (fourier.py:136-142) for a stack of coefficient vectors s [M, N] at scattered points -> DEVICE [M, npts]."""
        s_d = self.to_dev(np.asarray(s, dtype=np.complex128), _c128) if not torch.is_tensor(s) else s.to(self.device, _c128)
        s_d = s_d.reshape(-1, s_d.shape[-1]).contiguous()
        M, N = s_d.shape
        kx_d = self.to_dev(np.asarray(kx, dtype=np.complex128).reshape(-1), _c128)
        ky_d = self.to_dev(np.asarray(ky, dtype=np.complex128).reshape(-1), _c128)
        x_d = self.to_dev(np.asarray(x, dtype=np.float64).reshape(-1), _f64)
        y_d = self.to_dev(np.asarray(y, dtype=np.float64).reshape(-1), _f64)
        npts = x_d.numel()
        assert kx_d.numel() == N and ky_d.numel() == N and y_d.numel() == npts
        out = torch.empty((M, npts), dtype=_c128, device=self.device)
        cap = max(1, int((2 << 30) // (N * 16)))                     # points per call: phase matrix <= 2 GB
        for lo in range(0, npts, cap):
            hi = min(npts, lo + cap)
            part = out if (lo == 0 and hi == npts) else torch.empty((M, hi - lo), dtype=_c128, device=self.device)
            wb = self.lib.kh_idft_work_bytes(N, hi - lo)
            ws = self.workspace(wb)
            check(self.lib, self.lib.kh_idft_batch(M, N, hi - lo, _ptr(kx_d), _ptr(ky_d), _ptr(x_d[lo:hi]), _ptr(y_d[lo:hi]), _ptr(s_d), _ptr(part),
                                                   _ptr(ws), ws.numel(), self.stream()), "kh_idft_batch")
            if part is not out:
                out[:, lo:hi] = part
        return out

    def beam_amplitudes(self, kps, g, x, y, fields4, scale):
        """beams.amplitudes_from_fields for a batch of k-points -> DEVICE [B, N, 4] (Ex, Ey, Hx, Hy per harmonic)."""
        kp_d = self.to_dev(np.asarray(kps, dtype=np.complex128).reshape(-1, 2), _c128)
        g_d = self.to_dev(np.ascontiguousarray(g, dtype=np.float64), _f64)
        x_d = self.to_dev(np.asarray(x, dtype=np.float64).reshape(-1), _f64)
        y_d = self.to_dev(np.asarray(y, dtype=np.float64).reshape(-1), _f64)
        f_d = self.to_dev(np.asarray(fields4, dtype=np.complex128).reshape(-1, 4), _c128)
        B, N, npts = kp_d.shape[0], g_d.shape[1], x_d.numel()
        assert f_d.shape[0] == npts and y_d.numel() == npts
        amp = torch.empty((B, N, 4), dtype=_c128, device=self.device)
        cap = max(1, int((4 << 30) // (N * npts * 16)))               # k-points per call: phase matrix <= 4 GB
        for lo in range(0, B, cap):
            hi = min(B, lo + cap)
            wb = self.lib.kh_beam_amplitudes_work_bytes(hi - lo, N, npts)
            ws = self.workspace(wb)
            check(self.lib, self.lib.kh_beam_amplitudes(hi - lo, N, npts, _ptr(kp_d[lo:hi]), _ptr(g_d), _ptr(x_d), _ptr(y_d), _ptr(f_d), float(scale),
                                                        _ptr(amp[lo:hi]), _ptr(ws), ws.numel(), self.stream()), "kh_beam_amplitudes")
        return amp

    def _fields_grid(self, plan, solved, wl, kp, inc, xs, ys, z, stack_positions):
        B = solved["prefix"].shape[0]
        wl_d = self.to_dev(np.asarray(wl, dtype=np.float64).reshape(B), _f64)
        kp_d = self.to_dev(np.asarray(kp, dtype=np.complex128).reshape(B, 2), _c128)
        inc_d = self.to_dev(np.asarray(inc, dtype=np.complex128).reshape(B, 2, plan.n), _c128)
        xs_d = self.to_dev(np.asarray(xs, dtype=np.float64).reshape(-1), _f64)
        ys_d = self.to_dev(np.asarray(ys, dtype=np.float64).reshape(-1), _f64)
        z_h = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(-1))
        zp_h = np.ascontiguousarray(np.asarray(stack_positions, dtype=np.float64).reshape(-1))
        assert zp_h.size == plan.Ls + 1, "stack_positions must hold Ls+1 interface positions"
        nx, ny, nz = xs_d.numel(), ys_d.numel(), z_h.size
        F = torch.empty((B, nz, 6, ny * nx), dtype=_c128, device=self.device)
        out = Outputs()
        out.prefix_dev, out.suffix_dev = solved["prefix"].data_ptr(), solved["suffix"].data_ptr()
        out.W_dev, out.V_dev, out.L_dev = solved["W"].data_ptr(), solved["V"].data_ptr(), solved["L"].data_ptr()
        wb = self.lib.kh_fields_grid_workspace_bytes(plan.handle, B, nx, ny, nz)
        ws = self.workspace(wb)
        dp = C.POINTER(C.c_double)
        check(self.lib, self.lib.kh_fields_grid_batch(plan.handle, B, _ptr(wl_d), _ptr(kp_d), _ptr(inc_d), C.byref(out), _ptr(xs_d), nx, _ptr(ys_d), ny,
                                                      z_h.ctypes.data_as(dp), nz, zp_h.ctypes.data_as(dp), _ptr(F), _ptr(ws), ws.numel(), self.stream()),
              "kh_fields_grid_batch")
        return F
