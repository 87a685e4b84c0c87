"""Analytic Fourier coefficients of island shapes (host side, numpy + scipy.special) for
Crystal.add_layer_analytical / Layer.analytical.  Restates khepri/fourier.py:8-58, 145-159: the coefficients are
a few hundred numbers per layer, computed once per geometry; the convolution matrix is then built on the device by
the bit-exact Toeplitz gather (kh_toeplitz_gather) and inverted by kh_zinv_batched, exactly like a pixmap layer."""
import numpy as np


def fexpz(z):
    """(exp(z) - 1) / z with the reference's series below |z| = 1e-2 (fourier.py:8-20)."""
    z = np.asarray(z, dtype=complex)
    small = np.abs(z) <= 1e-2
    out = np.empty_like(z)
    zb = z[~small]
    out[~small] = (np.exp(zb) - 1.0) / zb
    zs = z[small]
    out[small] = 1 + zs / 2.0 * (1 + zs / 3.0 * (1 + zs / 4.0 * (1 + zs / 5.0 * (1 + zs / 6.0 * (1 + zs / 7)))))
    return out


def transform_rectangle(ll0, ll1, ur0, ur1, Gx, Gy, sigma):
    """fourier.py:28-40."""
    a, b = ur0 - ll0, ur1 - ll1
    return a * b / sigma * fexpz(-1j * Gx * a) * fexpz(-1j * Gy * b) * np.exp(-1j * (ll0 * Gx + ll1 * Gy))


def transform_disc(center0, center1, radius, Gx, Gy, sigma):
    """fourier.py:43-58 (Bessel J1; the G = 0 term is the filling fraction)."""
    from scipy import special
    norm = np.sqrt(Gx * Gx + Gy * Gy) * radius
    zero = np.logical_and(np.isclose(Gx, 0.0), np.isclose(Gy, 0.0))
    nz = ~zero
    out = np.zeros_like(Gx, dtype=complex)
    out[nz] = (np.pi * radius ** 2 / sigma * 2 * np.exp(-1j * (center0 * Gx[nz] + center1 * Gy[nz]))
               * special.jv(1.0, norm[nz]) / norm[nz])
    out[zero] = np.pi * radius ** 2 / sigma
    return out


_TRANSFORMS = {"rectangle": transform_rectangle, "disc": transform_disc}


def transform(shape, params, Gx, Gy, sigma):
    """fourier.py:23-25; shapes without a closed form raise like the reference's getattr would."""
    if shape not in _TRANSFORMS:
        raise AttributeError(f"no analytic Fourier transform for island type {shape!r}")
    return _TRANSFORMS[shape](*params, Gx, Gy, sigma)


def combine_fourier_masks(islands_data, eps_host, inverse=False):
    """fourier.py:145-159: eps_g = host * delta + sum_islands (eps_island - eps_host) * shape_g (or of 1/eps)."""
    length = islands_data[0][0].shape[0]
    center = (length - 1) // 2
    f = (lambda x: 1 / x) if inverse else (lambda x: x)
    eps_g = np.zeros((length,), dtype=complex)
    eps_g[center] = f(eps_host)
    for shape_g, eps_island in islands_data:
        eps_g += (f(eps_island) - f(eps_host)) * shape_g
    return eps_g


def analytical_coefficients(expansion, islands, eps_host):
    """Fourier coefficient table handed to the Toeplitz gather (layer.py:161-168): the islands' coefficients on the
    3x oversampled harmonic grid, flat index g = q * 3P + p, reshaped to (3P, 3Q) as the reference does (its `.T` acts
    on a 1-D array and is a no-op, so for P = Q the table is indexed [q, p]: kept for parity)."""
    Gx, Gy, epw = expansion.g_vectors_expanded(3)
    data = [(transform(isl["type"], isl["params"], Gx, Gy, expansion.sigma), isl["epsilon"]) for isl in islands]
    return combine_fourier_masks(data, eps_host, inverse=False).reshape(epw)


def idft(ffield, kx, ky, x, y, engine=None):
    """Fourier -> real space at the points (x, y) (fourier.py:136-142): f = sum_g ffield[g] exp(i (kx[g] x + ky[g] y)).
    Same arguments and result as the reference (ffield (N,) -> array of x's shape); a stack ffield (M, N) gives (M,) + x.shape
    in one launch pair (kh_idft_batch: phase matrix + one DMMA GEMM).  Runs on the engine's device; no CPU path."""
    from .engine import Engine
    eng = engine or Engine.default()
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    s = np.asarray(ffield, dtype=np.complex128)
    out = eng.idft(s.reshape(-1, s.shape[-1]), kx, ky, x, y).cpu().numpy()
    return out.reshape(s.shape[:-1] + x.shape)
