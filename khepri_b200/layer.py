"""Layer descriptors (host side).  Same constructors and attributes as khepri/layer.py:63-143; the
arithmetic of Layer.solve (layer.py:145-194) runs batched on the GPU through the Crystal."""
from enum import IntEnum

import numpy as np


class Formulation(IntEnum):          # layer.py:19-24
    UNIFORM = 0
    FFT = 1
    ANALYTICAL = 2
    HALF_SPACE_INC = 3
    HALF_SPACE_TRN = 4


class Field(IntEnum):                # layer.py:27-32
    X = 0
    Y = 1
    Z = 2
    NORM = 3
    POYNTING = 4


_FPW = {}


def _fingerprint_weights(size):
    w = _FPW.get(size)
    if w is None:
        rng = np.random.default_rng(0x5EED + size)
        w = _FPW[size] = (rng.standard_normal(size), rng.standard_normal(size))
    return w


class Layer:
    def __init__(self):
        self.formulation = None
        self.expansion = None
        self.W = None
        self.V = None
        self.L = None
        self.S = None
        self.IC = None
        self.fields = False
        self._cache = None           # device copies of (C, C^-1), keyed on the CONTENT of the pixmap / island list

    def content_key(self):
        """Hashable token of everything the layer's S-matrix depends on besides the source.  Keyed on content, not on
        object identity: the reference recomputes convolution_matrix(self.epsilon) on every solve (layer.py:157), so a
        pixmap mutated in place must invalidate the cached convolution matrix and the plan."""
        eps = self.epsilon
        if self.formulation == Formulation.FFT:
            # fingerprint of the pixel values: two weighted sums (a fixed pseudo-random weight vector per size).  Any in-place edit
            # changes them; it costs ~10 us per pixmap where a cryptographic hash of the bytes cost 0.35 ms per solve.
            arr = np.asarray(eps)
            flat = arr.reshape(-1)
            wgt = _fingerprint_weights(flat.size)
            tok = (arr.shape, str(arr.dtype), complex(flat.sum()), complex(np.dot(flat, wgt[0])), complex(np.dot(flat * flat, wgt[1])))
        elif self.formulation == Formulation.ANALYTICAL:
            tok = (repr(eps), complex(self.eps_host))
        else:
            tok = complex(eps)
        return (int(self.formulation), float(self.depth), tok)

    def eps_bound(self):
        """max |epsilon| of a patterned layer (bounds the spectrum of Omega^2 for the doubling method)."""
        if self.formulation == Formulation.FFT:
            return float(np.max(np.abs(self.epsilon)))
        if self.formulation == Formulation.ANALYTICAL:
            vals = [abs(complex(self.eps_host))] + [abs(complex(isl["epsilon"])) for isl in self.epsilon]
            return float(max(vals))
        return float(abs(complex(self.epsilon)))

    def eps_min_real(self):
        """min Re(epsilon) of a patterned layer.  While it stays well above zero the convolution matrix has its field of values in
        the right half plane and max |epsilon| bounds the spectrum of Omega^2; a metallic inclusion (Re epsilon < 0) does not."""
        if self.formulation == Formulation.FFT:
            return float(np.min(np.real(self.epsilon)))
        if self.formulation == Formulation.ANALYTICAL:
            return float(min([complex(self.eps_host).real] + [complex(isl["epsilon"]).real for isl in self.epsilon]))
        return float(complex(self.epsilon).real)

    @classmethod
    def pixmap_or_uniform(cls, expansion, pixmap, depth):
        eps0 = pixmap.flatten()[0]
        if np.all(pixmap == eps0):
            return cls.uniform(expansion, eps0, depth)
        return cls.pixmap(expansion, pixmap, depth)

    @classmethod
    def pixmap(cls, expansion, pixmap, depth):
        layer = cls()
        layer.expansion = expansion
        layer.formulation = Formulation.FFT
        layer.epsilon = pixmap
        layer.depth = depth
        return layer

    @classmethod
    def uniform(cls, expansion, epsilon, depth):
        layer = cls()
        layer.expansion = expansion
        layer.formulation = Formulation.UNIFORM
        layer.epsilon = epsilon
        layer.depth = depth
        return layer

    @classmethod
    def analytical(cls, expansion, islands_description, eps_host, depth):
        """layer.py:121-129: islands with closed-form Fourier transforms (Drawing.islands()) in a host medium."""
        layer = cls()
        layer.expansion = expansion
        layer.formulation = Formulation.ANALYTICAL
        layer.epsilon = islands_description
        layer.eps_host = eps_host
        layer.depth = depth
        return layer

    @classmethod
    def half_infinite(cls, expansion, type, epsilon):
        layer = cls()
        if type == "reflexion":
            layer.formulation = Formulation.HALF_SPACE_INC
        elif type == "transmission":
            layer.formulation = Formulation.HALF_SPACE_TRN
        else:
            raise ValueError("half_infinite type must be 'reflexion' or 'transmission'")
        layer.expansion = expansion
        layer.depth = 0
        layer.epsilon = epsilon
        return layer

    # -- device-side convolution matrix (tools.convolution_matrix + np.linalg.inv, layer.py:157-158),
    #    computed once per pixmap instead of once per solve
    def convmat_device(self, engine):
        key = (self.content_key(), tuple(self.expansion.pw), id(engine))
        if self._cache is None or self._cache[0] != key:
            if self.formulation == Formulation.ANALYTICAL:         # layer.py:161-168: analytic coefficients -> Toeplitz gather
                from .fourier import analytical_coefficients
                Cm = engine.toeplitz_gather(analytical_coefficients(self.expansion, self.epsilon, self.eps_host), self.expansion.pw)
            else:
                Cm = engine.convmat(np.asarray(self.epsilon), self.expansion.pw)[0]
            ICm, info = engine.zinv(Cm, return_info=True)
            if int(info.max().item()) != 0:
                raise np.linalg.LinAlgError("Singular matrix")     # what np.linalg.inv raises in the reference
            self._cache = (key, Cm, ICm)
        return self._cache[1], self._cache[2]

    # -- Layer.solve of the reference (layer.py:145-194): S, and with `fields` also W, V, L, IC, for ONE source.
    #    A one-layer, half-space-free Crystal through the same batched kernels (the star product with the identity is exact).
    def solve(self, k_parallel, wavelength, engine=None):
        from .crystal import Crystal
        cl = Crystal(self.expansion.pw, void=True, engine=engine)
        cl.expansion = self.expansion
        keep = self.fields
        cl.layers["L"] = self
        cl.set_device(["L"], [keep])
        cl.set_source(wavelength, kp=tuple(k_parallel))
        cl.solve()
        self.S = np.asarray(cl.Stot)
        self.fields = keep
        if self.formulation in (Formulation.FFT, Formulation.ANALYTICAL):
            Cm, ICm = self.convmat_device(cl.engine)
            self.C = Cm.cpu().numpy()
            self.IC = ICm.cpu().numpy() if keep else None
        else:
            self.IC = 1 / self.epsilon if keep else None
        if not keep:
            self.W = self.V = self.L = None
        return self


def stack_layers(pw, layers, mask, engine=None):
    """layer.py:35-60: forward partial products (kept where mask), reverse partial products, total, from solved layers."""
    from .alternative import redheffer_product, scattering_identity
    Stot = scattering_identity(pw, block=True)
    Sls = []
    for i, layer in enumerate(layers):
        Stot = redheffer_product(Stot, layer.S, engine=engine)
        Sls.append(Stot.copy() if mask[i] else None)
    Srev = scattering_identity(pw, block=True)
    rmask = list(reversed(mask[1:]))
    Srs = []
    for i, layer in enumerate(reversed(layers[1:])):
        Srs.append(Srev.copy() if rmask[i] else None)
        Srev = redheffer_product(np.asarray(layer.S).copy(), Srev, engine=engine)
    Srs.append(Srev.copy())
    return Sls, list(reversed(Srs)), Stot
