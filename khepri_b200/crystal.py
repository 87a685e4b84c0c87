"""Drop-in ``Crystal`` for the sweep loops of Kaeryv/Khepri (khepri/crystal.py:34-406).

Same constructor, mutators, compute calls, result attributes and error behaviour as the reference;
the arithmetic behind ``solve`` / ``poynting_flux_end`` / ``fields_volume`` runs in the batched
sm_100a kernels.  Added for throughput: ``solve_batch`` / ``sweep`` take whole arrays of
(wavelength, k-point) sources -- the scalar ``solve()`` is a batch of one.
"""
import logging
from cmath import sqrt as csqrt
from copy import copy
from types import SimpleNamespace

import numpy as np

from . import _lib
from .alternative import incident
from .engine import Engine, Plan
from .expansion import Expansion
from .extension import ExtendedLayer as EL
from .layer import Formulation, Layer
from .misc import ensure_array
from .tools import compute_kplanar


class Crystal:
    def __init__(self, pw, lattice="square", lattice_pitch=1, void=False, epsi=1, epse=1, engine=None):
        self.pw = pw
        self.a = lattice_pitch
        self.void = void
        self.epsi = epsi
        self.epse = epse
        self.source = None
        if isinstance(lattice, str):
            if lattice == "square":
                self.lattice = self.a * np.asarray([[1, 0], [0, 1]])
            elif lattice == "hexagonal":
                self.lattice = self.a * np.asarray([[np.sqrt(3) / 2, 0.5], [np.sqrt(3) / 2, -0.5]])
            else:
                raise NotImplementedError(f"This {lattice} magic-string is not emplemented.")
        else:
            self.lattice = lattice
        self.expansion = Expansion(pw, self.lattice)
        self.layers = dict()
        self.stacking_matrices = list()
        self.stacking_reverse_matrices = list()
        self.stack_positions = []
        self._engine = engine
        # how patterned layers get their S-matrix: "auto" | "eig" | "doubling" (Engine._select_method); results agree to rounding
        self.method = "auto"
        self._plan = None
        self._plan_key = None
        self._S_host = None
        self._S_dev = None
        self._solved = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_expansion(cls, expansion, **kwargs):
        obj = cls(expansion.pw, **kwargs)
        obj.expansion = expansion
        return obj

    def add_layer_uniform(self, name, epsilon, depth):
        self.layers[name] = Layer.uniform(self.expansion, epsilon, depth)

    def add_pixmap_or_uniform(self, name, epsilon, depth):
        sample = epsilon.flatten()[0]
        if np.all(epsilon == sample):
            self.add_layer_uniform(name, sample, depth)
        else:
            self.add_layer_pixmap(name, epsilon, depth)

    def add_layer_pixmap(self, name, epsilon, depth):
        self.layers[name] = Layer.pixmap(self.expansion, epsilon, depth)

    def add_layer_analytical(self, name, epsilon, epsilon_host, depth):
        self.layers[name] = Layer.analytical(self.expansion, epsilon, epsilon_host, depth)

    def add_layer(self, name, layer, extended=False):
        self.layers[name] = EL(self.expansion, layer) if extended else layer

    def set_device(self, layers_stack, fields_mask=False):
        """crystal.py:131-164: pre/append the incidence and emergence half spaces."""
        self.device_stack = copy(layers_stack)
        self.global_stacking = []
        if not self.void:
            self.global_stacking.append("Sref")
        self.global_stacking.extend(layers_stack)
        if not self.void:
            self.global_stacking.append("Strans")
        required = set(self.global_stacking)
        if "Sref" in required and "Sref" not in self.layers:
            self.layers["Sref"] = Layer.half_infinite(self.expansion, "reflexion", self.epsi)
            self.layers["Sref"].fields = True
        if "Strans" in required and "Strans" not in self.layers:
            self.layers["Strans"] = Layer.half_infinite(self.expansion, "transmission", self.epse)
            self.layers["Strans"].fields = True
        fields_mask = [fields_mask] * len(layers_stack) if isinstance(fields_mask, bool) else fields_mask
        self.stack_retain_mask = [True]
        self.stack_retain_mask.extend(fields_mask)
        self.stack_retain_mask.append(True)
        # (void stacks have no half spaces: align the mask with the stack -- the reference zips the padded mask, crystal.py:157-159)
        aligned = self.stack_retain_mask if not self.void else self.stack_retain_mask[1:-1]
        for name, enabled in zip(self.global_stacking, aligned):
            self.layers[name].fields |= enabled
            if hasattr(self.layers[name], "base"):
                self.layers[name].base.fields |= enabled
        self._plan = None

    # ------------------------------------------------------------------ properties
    @property
    def engine(self):
        if self._engine is None:
            self._engine = Engine.default()
        return self._engine

    @property
    def depth(self):
        return sum(self.layers[name].depth for name in self.device_stack)

    @property
    def source_defined(self):
        return self.source is not None

    @property
    def solved(self):
        return self._S_host is not None or self._S_dev is not None

    @property
    def zmax(self):
        return self.stack_positions[-2]

    @property
    def Stot(self):
        if self._S_host is None and self._S_dev is not None:
            self._S_host = self._S_dev.cpu().numpy()
        return self._S_host

    @Stot.setter
    def Stot(self, value):
        # callers do e.g. `cl.Stot = redheffer_product(cl.Stot, cl.Stot)` (examples/crystal_api/woodpile.py:85)
        self._S_host = None if value is None else np.asarray(value, dtype=np.complex128)
        self._S_dev = None

    S = Stot

    # ------------------------------------------------------------------ plan (geometry -> device)
    def _build_plan(self, want_fields):
        eng = self.engine
        names, table = [], []

        def index_of(name):
            return names.index(name)

        ext_seen = False
        for name in self.global_stacking:
            if name in names:
                continue
            layer = self.layers[name]
            if isinstance(layer, EL):
                ext_seen = True
                base = layer.base
                bname = ("__base__", name)
                names.append(bname)
                table.append(self._layer_entry(base, eng, False))
                names.append(name)
                table.append({"kind": _lib.LAYER_EXTENDED, "depth": layer.depth, "ext_base": index_of(bname), "ext_mode": layer.mode})
            else:
                names.append(name)
                table.append(self._layer_entry(layer, eng, want_fields and layer.fields))
        stack = [index_of(name) for name in self.global_stacking]
        Nb, glhs, grhs = 0, None, None
        if ext_seen:
            lhs, rhs = self.expansion.expansion_lhs, self.expansion.expansion_rhs
            Nb = int(np.prod(lhs.pw))
            glhs, grhs = lhs._g_vectors, rhs._g_vectors
        pw = tuple(int(p) for p in self.expansion.pw)
        plan = Plan(eng, pw, self.expansion._g_vectors, self.epsi, self.epse, table, stack, Nb, glhs, grhs)
        plan.names = names
        return plan

    @staticmethod
    def _layer_entry(layer, eng, retain):
        f = layer.formulation
        if f == Formulation.UNIFORM:
            return {"kind": _lib.LAYER_UNIFORM, "eps": layer.epsilon, "depth": layer.depth, "retain": retain}
        if f == Formulation.FFT or f == Formulation.ANALYTICAL:
            Cm, ICm = layer.convmat_device(eng)
            return {"kind": _lib.LAYER_PIXMAP, "depth": layer.depth, "C": Cm, "IC": ICm, "retain": retain, "eps_bound": layer.eps_bound(), "eps_min_real": layer.eps_min_real()}
        if f == Formulation.HALF_SPACE_INC:
            return {"kind": _lib.LAYER_HALF_INC, "eps": layer.epsilon, "retain": retain}
        if f == Formulation.HALF_SPACE_TRN:
            return {"kind": _lib.LAYER_HALF_TRN, "eps": layer.epsilon, "retain": retain}
        raise NotImplementedError(f"formulation {f} is not supported on the GPU path yet")

    def _geometry_key(self, want_fields):
        key = [want_fields, id(self.expansion), complex(self.epsi), complex(self.epse), tuple(self.global_stacking)]
        for name in dict.fromkeys(self.global_stacking):            # each distinct layer once
            layer = self.layers[name]
            base = layer.base if isinstance(layer, EL) else layer
            key.append((name, id(layer), base.content_key()))
        key.append(self.expansion._g_vectors.tobytes())
        return tuple(key)

    def _get_plan(self, want_fields=False):
        key = self._geometry_key(want_fields)
        if self._plan is None or self._plan_key != key:
            self._plan = self._build_plan(want_fields)
            self._plan_key = key
        return self._plan

    # ------------------------------------------------------------------ source / solve
    def set_source(self, wavelength, te=1.0, tm=1.0, theta=0.0, phi=0.0, kp=None):
        self.source = SimpleNamespace(te=te, tm=tm, theta=theta, phi=phi, wavelength=wavelength)
        if kp is not None:
            self.kp = kxi, kyi = kp
        else:
            self.kp = kxi, kyi = compute_kplanar(self.epsi, wavelength, theta, phi)
        self.k0 = 2 * np.pi / wavelength
        self.kzi = np.conj(csqrt(self.k0 ** 2 * self.epsi - kxi ** 2 - kyi ** 2))

    def _wants_fields(self):
        return any(self.layers[name].fields for name in self.device_stack)

    def solve(self):
        assert self.source_defined, "Call set_source before solving."
        logging.debug("Solving each required layer")
        want_fields = self._wants_fields()
        plan = self._get_plan(want_fields)
        res = self.engine.solve_batch(plan, [self.source.wavelength], [self.kp], want_S=True, want_flux=False, want_fields=want_fields, method=self.method)
        self._check_info(res["info"])
        layer_sizes = [self.layers[name].depth for name in self.global_stacking]
        self.stack_positions = list(np.cumsum(layer_sizes))
        self.stack_positions[-1] = np.inf
        if not self.void:
            self.stack_positions.insert(0, -np.inf)
        self._S_dev = res["Stot"][0]
        self._S_host = None
        self._solved = res if want_fields else None
        self._solved_src = (self.source.wavelength, tuple(self.kp))
        if want_fields:
            pre, suf = res["prefix"][0].cpu().numpy(), res["suffix"][0].cpu().numpy()
            mask = self.stack_retain_mask if not self.void else self.stack_retain_mask[1:-1]
            self.stacking_matrices = [pre[i] if mask[i] else None for i in range(len(mask))]
            self.stacking_reverse_matrices = [suf[i] if (i == 0 or mask[i]) else None for i in range(len(mask))]
            W, V, L = res["W"][0].cpu().numpy(), res["V"][0].cpu().numpy(), res["L"][0].cpu().numpy()
            for idx, name in enumerate(plan.names):
                layer = self.layers.get(name) if isinstance(name, str) else None
                if layer is not None and layer.fields:
                    layer.W, layer.V, layer.L = W[idx], V[idx], L[idx]
        else:
            self.stacking_matrices, self.stacking_reverse_matrices = [], []

    @staticmethod
    def _check_info(info):
        bad = int(info.max().item()) if info.numel() else 0
        if bad & 2:
            raise np.linalg.LinAlgError("Singular matrix")
        if bad & 4:
            raise np.linalg.LinAlgError("doubling method: spectral bound exceeded (use method='eig')")
        if bad & 8:          # (only with method='doubling' forced: 'auto' has re-solved such sources with the eigen method)
            raise np.linalg.LinAlgError("doubling method: ill-conditioned self star product at a sub-slab resonance (use method='auto' or 'eig')")
        if bad & 1:
            raise np.linalg.LinAlgError("Eigenvalues did not converge")

    def _S_device(self):
        if self._S_dev is None:
            self._S_dev = self.engine.to_dev(self._S_host, __import__("torch").complex128)
        return self._S_dev

    def poynting_flux_end(self, only_total=True):
        assert self.solved, "Call solve first"
        plan = self._plan if self._plan is not None else self._get_plan(False)      # the plan the last solve() used
        pol = [(self.source.te, self.source.tm)]
        out = self.engine.flux(plan, self._S_device(), [self.source.wavelength], [self.kp], pol, want_orders=not only_total)
        if only_total:
            R, T = out[0].cpu().numpy()
            return float(R), float(T)
        RT, orders = out
        RT, orders = RT[0].cpu().numpy(), orders[0].cpu().numpy()
        return (RT[0], orders[0]), (RT[1], orders[1])

    # ------------------------------------------------------------------ batched sweeps (new)
    def solve_batch(self, wavelengths, kps=None, te=1.0, tm=1.0, theta=0.0, phi=0.0, only_total=True, return_S=False, chunk=None):
        """R, T for arrays of sources: one entry = set_source(...); solve(); poynting_flux_end().

        wavelengths [B]; either kps [B, 2] or theta/phi (degrees, scalar or [B]); te/tm scalar or [B].
        Returns (R[B], T[B]) numpy arrays (plus per-order arrays [B, N] when only_total is False,
        plus Stot [B,2,2,n,n] when return_S).
        """
        wl = np.atleast_1d(np.asarray(wavelengths, dtype=np.float64)).reshape(-1)
        B = wl.size
        if kps is None:
            th = np.deg2rad(np.broadcast_to(np.asarray(theta, dtype=np.float64), (B,)))
            ph = np.deg2rad(np.broadcast_to(np.asarray(phi, dtype=np.float64), (B,)))
            amp = csqrt(self.epsi) * 2 * np.pi / wl * np.sin(th)
            kp = np.stack([np.cos(ph).astype(complex) * amp, np.sin(ph).astype(complex) * amp], axis=1)
        else:
            kp = np.asarray(kps, dtype=np.complex128).reshape(B, 2)
        pol = np.stack([np.broadcast_to(np.asarray(te, dtype=np.complex128), (B,)),
                        np.broadcast_to(np.asarray(tm, dtype=np.complex128), (B,))], axis=1)
        plan = self._get_plan(False)
        res = self.engine.solve_batch(plan, wl, kp, pol, want_S=return_S, want_flux=True, want_orders=not only_total, chunk=chunk, method=self.method)
        self._check_info(res["info"])
        RT = res["RT"].cpu().numpy()
        out = [RT[:, 0], RT[:, 1]]
        if not only_total:
            orders = res["orders"].cpu().numpy()
            out = [(RT[:, 0], orders[:, 0]), (RT[:, 1], orders[:, 1])]
        if return_S:
            out.append(res["Stot"].cpu().numpy())
        return tuple(out)

    sweep = solve_batch

    # ------------------------------------------------------------------ fields
    def locate_layer(self, z):
        """crystal.py:208-232."""
        assert z != np.nan
        layer_index = int(np.searchsorted(self.stack_positions, z) - 1)
        layer_name = self.global_stacking[layer_index]
        zr = z if z <= 0 else z - self.stack_positions[layer_index]
        return self.layers[layer_name], layer_index, zr

    def get_source_as_field_vectors(self):
        efield = incident(self.pw, self.source.te, self.source.tm, k_vector=(self.kp[0], self.kp[1], self.kzi))
        return efield, np.zeros_like(efield)

    def _fields_points(self, x, y, zs, incident_fields):
        assert self.solved, "Call solve first."
        if self._solved is None or self._solved_src != (self.source.wavelength, tuple(self.kp)):
            raise AssertionError("Layer at z did not store eigenspace.")     # crystal.py:245
        for z in zs:
            layer, _, _ = self.locate_layer(z)
            assert layer.fields, f"Layer at {z} did not store eigenspace."
        plan = self._get_plan(True)
        inc = np.asarray(incident_fields, dtype=np.complex128).reshape(1, 2, plan.n)
        grid = None
        if x.ndim == 2 and x.shape == y.shape and np.all(x == x[:1, :]) and np.all(y == y[:, :1]):
            grid = (x[0, :], y[:, 0])                       # numpy.meshgrid(indexing="xy") coordinates: separable transform
        F = self.engine.fields(plan, self._solved, [self.source.wavelength], [self.kp], inc, x.ravel(), y.ravel(), zs,
                               self.stack_positions, grid=grid)
        return F[0]

    def _fields_fourier(self, zs, incident_fields):
        """crystal.py:234-277 for a list of depths -> DEVICE [nz, 6, N]."""
        assert self.solved, "Call solve first."
        if self._solved is None or self._solved_src != (self.source.wavelength, tuple(self.kp)):
            raise AssertionError("Layer at z did not store eigenspace.")     # crystal.py:245
        for z in zs:
            layer, _, _ = self.locate_layer(z)
            assert layer.fields, f"Layer at {z} did not store eigenspace."
        plan = self._get_plan(True)
        inc = np.asarray(incident_fields, dtype=np.complex128).reshape(1, 2, plan.n)
        return self.engine.fields_fourier(plan, self._solved, [self.source.wavelength], [self.kp], inc, zs, self.stack_positions)[0]

    def _fourier_fields(self, z, incident_fields):
        """crystal.py:234-277 -> (sx, sy, sz, ux, uy, uz) Fourier vectors at depth z."""
        if isinstance(incident_fields, tuple) and len(incident_fields) == 2:
            incident_fields = np.hstack(incident_fields)
        return tuple(self._fields_fourier([float(z)], incident_fields).cpu().numpy()[0])

    def fields_batch_sum(self, wavelengths, kps, incident_fields, x, y, zs, chunk=64):
        """Sum over a batch of sources of the field maps (Ex,Ey,Ez,Hx,Hy,Hz)(z, y, x): the k-sum of the Brillouin-zone
        integration loop (examples/bzi/bzi_animation.py:59-80) -- solve with retained eigenspaces, reconstruct with each
        source's own incident amplitudes [B, 4N] = (Ex_g, Ey_g, Hx_g, Hy_g), accumulate on the device.
        Returns a DEVICE tensor [nz, 6, ny * nx]."""
        import torch
        wl = np.atleast_1d(np.asarray(wavelengths, dtype=np.float64)).reshape(-1)
        B = wl.size
        kp = np.asarray(kps, dtype=np.complex128).reshape(B, 2)
        plan = self._get_plan(True)
        inc = np.asarray(incident_fields, dtype=np.complex128).reshape(B, 2, plan.n)
        layer_sizes = [self.layers[name].depth for name in self.global_stacking]
        pos = list(np.cumsum(layer_sizes))
        pos[-1] = np.inf
        if not self.void:
            pos.insert(0, -np.inf)
        x, y = ensure_array(x), ensure_array(y)
        grid = None
        if x.ndim == 2 and x.shape == y.shape and np.all(x == x[:1, :]) and np.all(y == y[:, :1]):
            grid = (x[0, :], y[:, 0])
        total = torch.zeros((len(zs), 6, x.size), dtype=torch.complex128, device=self.engine.device)
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            res = self.engine.solve_batch(plan, wl[lo:hi], kp[lo:hi], want_S=False, want_flux=False, want_fields=True)
            self._check_info(res["info"])
            F = self.engine.fields(plan, res, wl[lo:hi], kp[lo:hi], inc[lo:hi], x.ravel(), y.ravel(), zs, pos, grid=grid)
            total += F.sum(dim=0)
        return total

    def fields_coords_xy(self, x, y, z, incident_fields=None, kp=None, return_fourier=False):
        """E and H at the points (x, y) of depth z (crystal.py:297-333) -> [E(3,ny,nx), H(3,ny,nx)]."""
        x, y = ensure_array(x), ensure_array(y)
        assert x.shape == y.shape and (len(x.shape) == 2), "x and y must be 2D meshgrids"
        assert self.solved, "Call solve first."
        if isinstance(z, str):
            raise NotImplementedError("'farfield' is not available (the reference's own far-field path calls an undefined function, crystal.py:323)")
        if incident_fields is None:
            incident_fields = np.hstack(self.get_source_as_field_vectors())
        elif isinstance(incident_fields, tuple) and len(incident_fields) == 2:
            incident_fields = np.hstack(incident_fields)
        if return_fourier:                                  # crystal.py:326-327 -> (sx, sy, sz, ux, uy, uz), each (N,)
            return self._fourier_fields(z, incident_fields)
        F = self._fields_points(x, y, [float(z)], incident_fields).cpu().numpy()[0]
        F = F.reshape((6,) + x.shape)
        return np.split(F, 2, axis=0)

    def fields_volume(self, x, y, z, incident_fields=None):
        """crystal.py:335-343 -> (E, H), each (nz, 3, ny, nx) complex128."""
        x, y = ensure_array(x), ensure_array(y)
        assert x.shape == y.shape and (len(x.shape) == 2), "x and y must be 2D meshgrids"
        if incident_fields is None:
            incident_fields = self.get_source_as_field_vectors()
        zs = [float(zi) for zi in z]
        F = self._fields_points(x, y, zs, np.hstack(incident_fields)).cpu().numpy()
        F = F.reshape((len(zs), 6) + x.shape)
        return F[:, :3], F[:, 3:]


class Multilayer(Crystal):
    def __init__(self, epsi=1, epse=1, engine=None):
        super().__init__((1, 1), lattice="square", lattice_pitch=1, void=False, epsi=epsi, epse=epse, engine=engine)
