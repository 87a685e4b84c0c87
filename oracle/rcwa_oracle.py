"""CPU oracle for the RCWA hot path (TEST INFRASTRUCTURE ONLY).

This module is a numpy/LAPACK *restatement* of the algorithm that the reference
(Kaeryv/Khepri, pure Python) runs behind ``khepri.crystal.Crystal``.  It is the
checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package ``khepri_b200`` never imports anything from ``oracle/``.

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the unmodified reference
(imported read-only from /root/reference in the build container) on the configs
of SURVEY.md §8(d) and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against every one of
them, plus the reference's own Fresnel known-answer test
(test/integration/test_complex_eps.py:14-42).

The arithmetic that matters lives in numpy (unpinned in the reference's
requirements.txt:1-5): ``numpy.linalg.solve/inv/eig`` (LAPACK zgesv/zgetri/zgeev)
and ``numpy.fft.fft2`` (pocketfft); the same calls are made here at the same
places so that the CPU timing is representative of the reference.

A *structure* is a plain dict::

    {"pw": (P, Q), "lattice": 2x2 array (rows = lattice vectors),
     "epsi": eps_incidence, "epse": eps_emergence,
     "layers": {name: ("uniform", eps, depth) | ("pixmap", eps_xy, depth) | ("analytical", islands, depth, eps_host, lattice)},
     "stack": [names...]}            # device stack WITHOUT the two half spaces

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import cmath
import math

import numpy as np
from numpy.lib.scimath import sqrt as csqrt
from numpy.linalg import eig, inv, solve

TWO_PI = 2.0 * math.pi


# --------------------------------------------------------------------------- #
# basis / geometry (khepri/expansion.py, khepri/tools.py)
# --------------------------------------------------------------------------- #
def harmonic_indices(pw):
    """Integer harmonic grid, x index fastest.  expansion.py:4-16."""
    P, Q = pw
    if P % 2 != 1 or Q % 2 != 1:
        raise AssertionError("pw entries must be odd")
    p = np.arange(P) - (P - 1) // 2
    q = np.arange(Q) - (Q - 1) // 2
    return np.stack([np.tile(p, Q), np.repeat(q, P)])


def reciprocal_basis(a1, a2):
    """tools.py:64-69."""
    f = TWO_PI / (a1[0] * a2[1] - a1[1] * a2[0])
    return (a2[1] * f, -a2[0] * f), (-a1[1] * f, a1[0] * f)


def g_vectors(pw, lattice):
    """expansion.py:30-41 -- g = m*b1 + n*b2, shape (2, N)."""
    b = np.asarray(reciprocal_basis(lattice[0], lattice[1]), dtype=float)
    idx = harmonic_indices(pw)
    return b[0][:, None] * idx[0][None, :] + b[1][:, None] * idx[1][None, :]


def k_vectors(g, kp, wl, eps=1):
    """Normalised (Kx, Ky, Kz).  expansion.py:19-26, 43-50."""
    k0 = TWO_PI / wl
    kx = np.asarray(kp[0]) + g[0].astype(complex)
    ky = np.asarray(kp[1]) + g[1].astype(complex)
    kz = np.conj(np.sqrt((k0 ** 2 * np.conj(eps) - kx ** 2 - ky ** 2).astype(complex)))
    return kx / k0, ky / k0, kz / k0


def kplanar(eps_inc, wl, theta_deg=0.0, phi_deg=0.0):
    """tools.py:17-21 (angles in degrees)."""
    th, ph = np.deg2rad(theta_deg), np.deg2rad(phi_deg)
    amp = cmath.sqrt(eps_inc) * TWO_PI / wl * np.sin(th)
    return np.array([np.cos(ph), np.sin(ph)], dtype=complex) * amp


def toeplitz_gather(F, pw):
    """tools.py:38-56: C[qr*P+pr, qc*P+pc] = F[Nx//2 + pr-pc, Ny//2 + qr-qc].

    Pure indexing (bit exact).  Negative indices wrap like numpy's.
    """
    P, Q = pw
    Nx, Ny = F.shape
    p = np.tile(np.arange(P), Q)
    q = np.repeat(np.arange(Q), P)
    ix = Nx // 2 + (p[:, None] - p[None, :])
    iy = Ny // 2 + (q[:, None] - q[None, :])
    if ix.max() >= Nx or iy.max() >= Ny or ix.min() < -Nx or iy.min() < -Ny:
        raise IndexError("harmonic differences exceed the Fourier grid")
    return np.asarray(F)[ix, iy].astype(complex)


def fourier_coefficients(pixmap):
    """tools.py:33-35: fftshift(fft2(eps)) / (Nx*Ny)."""
    pixmap = np.asarray(pixmap)
    return np.fft.fftshift(np.fft.fft2(pixmap)) / pixmap.size


def convolution_matrix(pixmap, pw):
    """tools.py:33-56."""
    return toeplitz_gather(fourier_coefficients(pixmap), pw)


def _fexpz(z):
    """fourier.py:8-20: (exp(z) - 1) / z, Horner series for |z| <= 1e-2."""
    z = np.asarray(z, dtype=complex)
    big = np.abs(z) > 1e-2
    r = np.empty_like(z)
    r[big] = (np.exp(z[big]) - 1.0) / z[big]
    zs = z[~big]
    r[~big] = 1 + zs / 2.0 * (1 + zs / 3.0 * (1 + zs / 4.0 * (1 + zs / 5.0 * (1 + zs / 6.0 * (1 + zs / 7)))))
    return r


def island_transform(kind, params, Gx, Gy, sigma):
    """fourier.py:23-58: analytic Fourier transform of a rectangle / disc island."""
    if kind == "rectangle":
        ll0, ll1, ur0, ur1 = params
        a, b = ur0 - ll0, ur1 - ll1
        return a * b / sigma * _fexpz(-1j * Gx * a) * _fexpz(-1j * Gy * b) * np.exp(-1j * (ll0 * Gx + ll1 * Gy))
    if kind == "disc":
        from scipy import special
        c0, c1, radius = params
        norm = np.sqrt(Gx * Gx + Gy * Gy) * radius
        zero = np.logical_and(np.isclose(Gx, 0.0), np.isclose(Gy, 0.0))
        out = np.zeros_like(Gx, dtype=complex)
        out[~zero] = (np.pi * radius ** 2 / sigma * 2 * np.exp(-1j * (c0 * Gx[~zero] + c1 * Gy[~zero]))
                      * special.jv(1.0, norm[~zero]) / norm[~zero])
        out[zero] = np.pi * radius ** 2 / sigma
        return out
    raise AttributeError(kind)


def analytical_convolution_matrix(islands, eps_host, pw, lattice):
    """layer.py:161-168 with expansion.py:86-92 and fourier.py:145-159: coefficients on the 3x oversampled harmonic
    grid (flat, p fastest), reshaped to (3P, 3Q) -- the reference's `.T` on the 1-D array is a no-op -- then gathered."""
    epw = [e * 3 if e > 1 else 1 for e in pw]
    b = np.asarray(reciprocal_basis(lattice[0], lattice[1]), dtype=float)
    idx = harmonic_indices(epw)
    G = b[0][:, None] * idx[0][None, :] + b[1][:, None] * idx[1][None, :]
    sigma = abs(lattice[0][0] * lattice[1][1] - lattice[0][1] * lattice[1][0])          # tools.unitcellarea
    eps_g = np.zeros(G.shape[1], dtype=complex)
    eps_g[(G.shape[1] - 1) // 2] = eps_host
    for isl in islands:
        eps_g += (isl["epsilon"] - eps_host) * island_transform(isl["type"], isl["params"], G[0], G[1], sigma)
    return toeplitz_gather(eps_g.reshape(epw), pw)


# --------------------------------------------------------------------------- #
# layer eigenmodes and S-matrices (khepri/alternative.py)
# --------------------------------------------------------------------------- #
def _q_blocks(Kx, Ky, eps):
    """[[KxKy, eps - Kx^2], [Ky^2 - eps, -KyKx]] with diagonal blocks."""
    N = len(Kx)
    Q = np.zeros((2 * N, 2 * N), dtype=complex)
    d = np.arange(N)
    Q[d, d] = Kx * Ky
    Q[d, N + d] = eps - Kx * Kx
    Q[N + d, d] = Ky * Ky - eps
    Q[N + d, N + d] = -Ky * Kx
    return Q


def _branch_kz(arg):
    """alternative.py:92-96 / 146-149: sign of Re(kz^2) picks the branch."""
    kz = np.array(arg, dtype=complex)
    neg = kz.real < 0
    kz[neg] = -1j * np.sqrt(-kz[neg])
    kz[~neg] = np.sqrt(kz[~neg])
    return kz


def free_space_modes(Kx, Ky):
    """alternative.py:84-99 -> (W0, V0)."""
    N = len(Kx)
    Q0 = _q_blocks(Kx, Ky, 1.0)
    kz = _branch_kz(1.0 - Kx * Kx - Ky * Ky)
    lam = np.concatenate([1j * kz, 1j * kz])
    return np.identity(2 * N), Q0 / lam


def uniform_layer_modes(Kx, Ky, eps):
    """alternative.py:130-156 -> (W, V, lambda)."""
    N = len(Kx)
    Qm = eps * ((1.0 / eps) * _q_blocks(Kx, Ky, eps))
    kz = _branch_kz(eps - Kx * Kx - Ky * Ky)
    lam = np.concatenate([1j * kz, 1j * kz])
    return np.identity(2 * N), Qm / lam, lam


def pq_matrices(Kx, Ky, C):
    """alternative.py:160-171: the coupled first-order system d/dz' [s; u] = [[0, P], [Q, 0]] [s; u] (Laurent rule)."""
    N = len(Kx)
    dKx, dKy = np.diag(Kx), np.diag(Ky)
    one = np.eye(N)
    iCKx, iCKy = solve(C, dKx), solve(C, dKy)
    Pm = np.block([[dKx @ iCKy, one - dKx @ iCKx],
                   [dKy @ iCKy - one, -dKy @ iCKx]]).astype(complex)
    Qm = np.block([[dKx @ dKy, C - dKx @ dKx],
                   [dKy @ dKy - C, -dKy @ dKx]]).astype(complex)
    return Pm, Qm


def structured_layer_modes(Kx, Ky, C):
    """alternative.py:158-178: Omega^2 = P Q, eig, lambda = sqrt, V = Q W / lambda."""
    Pm, Qm = pq_matrices(Kx, Ky, C)
    lam2, W = eig(Pm @ Qm)
    lam = np.sqrt(lam2 + 0j)
    return W, Qm @ W / lam, lam


def layer_smatrix_doubling(Pm, Qm, W0, V0, depth, k0, slicing_pow=3):
    """The reference's LEGACY layer algorithm (khepri/tmat/scattering.py:25-51) restated in the Crystal path's field basis:
    transfer matrix of a thin slice by scipy's expm (scattering.py:42-43), change to the mode basis of the free-space gaps
    (its U ... Vi, here R0 = [[W0, W0], [-V0, V0]], fields.py:46-51), matrix_s (tmat/matrices.py:167-176), then `slicing_pow`
    self star products (scattering.py:46-49, multS = redheffer_product).  No eigen-decomposition: an independent check of
    structured_layer_modes + layer_smatrix, and the oracle of the CUDA path's "doubling" method."""
    from scipy.linalg import expm
    n = Pm.shape[0]
    zero = np.zeros((n, n))
    gen = np.block([[zero, Pm], [Qm, zero]])
    M = expm(gen * (k0 * depth / 2 ** slicing_pow))
    R0 = np.block([[W0, W0], [-V0, V0]])
    T = solve(R0, M @ R0)                               # [c+; c-] at the right face = T [c+; c-] at the left face
    T11, T12, T21, T22 = T[:n, :n], T[:n, n:], T[n:, :n], T[n:, n:]
    S12 = inv(T22)
    S = np.array([[-S12 @ T21, S12], [T11 - T12 @ S12 @ T21, T12 @ S12]])
    for _ in range(slicing_pow):
        S = star(S, S)
    return S


def layer_smatrix(W, V, W0, V0, lam, depth, k0):
    """alternative.py:181-195 (symmetric slab between free-space gaps)."""
    a = solve(W, W0)
    b = solve(V, V0)
    A, B = a + b, a - b
    X = np.diag(np.exp(-lam * depth * k0))
    XB = X @ B
    T = A - XB @ solve(A, X) @ B
    S11 = solve(T, XB @ solve(A, X) @ A - B)
    S12 = solve(T, X @ (A - B @ solve(A, B)))
    return np.array([[S11, S12], [S12, S11]])


def _halfspace_modes(Kx, Ky, eps):
    N = len(Kx)
    Qh = _q_blocks(Kx, Ky, eps)
    kz = np.conj(csqrt((eps - Kx * Kx - Ky * Ky).astype(complex)))
    lam = np.concatenate([1j * kz, 1j * kz])
    return np.identity(2 * N), Qh / lam, lam


def halfspace_reflection(Kx, Ky, W0, V0, eps):
    """alternative.py:32-56 -> (S, W, V, lambda)."""
    W, V, lam = _halfspace_modes(Kx, Ky, eps)
    a, b = solve(W0, W), solve(V0, V)
    A, B = a + b, a - b
    iA = inv(A)
    AB = solve(A, B)
    S = np.array([[-AB, 2 * iA], [0.5 * (A - B @ AB), B @ iA]])
    return S, W, V, lam


def halfspace_transmission(Kx, Ky, W0, V0, eps):
    """alternative.py:59-82 -> (S, W, V, lambda)."""
    W, V, lam = _halfspace_modes(Kx, Ky, eps)
    a, b = solve(W0, W), solve(V0, V)
    A, B = a + b, a - b
    iA = inv(A)
    AB = solve(A, B)
    S = np.array([[B @ iA, 0.5 * (A - B @ AB)], [2 * iA, -AB]])
    return S, W, V, lam


def star(SA, SB):
    """Redheffer star product.  alternative.py:19-30."""
    n = SA.shape[-1]
    one = np.eye(n, dtype=complex)
    D = one - SB[0, 0] @ SA[1, 1]
    F = one - SA[1, 1] @ SB[0, 0]
    out = np.empty((2, 2, n, n), dtype=complex)
    out[0, 0] = SA[0, 0] + SA[0, 1] @ solve(D, SB[0, 0]) @ SA[1, 0]
    out[0, 1] = SA[0, 1] @ solve(D, SB[0, 1])
    out[1, 0] = SB[1, 0] @ solve(F, SA[1, 0])
    out[1, 1] = SB[1, 1] + SB[1, 0] @ solve(F, SA[1, 1]) @ SB[0, 1]
    return out


def identity_smatrix(n):
    """alternative.py:220-232 (block form)."""
    S = np.zeros((2, 2, n, n), dtype=complex)
    S[0, 1] = S[1, 0] = np.eye(n)
    return S


def incident_vector(pw, te, tm, kvec, normalize=True):
    """alternative.py:101-128: 2N source vector, delta at harmonic (N-1)//2."""
    if normalize:
        nrm = math.hypot(abs(te), abs(tm))
        te, tm = te / nrm, tm / nrm
    kvec = np.asarray(kvec, dtype=complex)
    kbar = kvec / np.linalg.norm(kvec)
    if abs(np.linalg.norm(kvec[:2])) < 1e-8:
        aTE = np.array([1, 0, 0], dtype=complex)
        aTM = np.array([0, 1, 0], dtype=complex)
    else:
        nz = np.array([0, 0, -1], dtype=complex)
        aTE = -np.cross(nz, kbar)
        aTE = aTE / np.linalg.norm(aTE)
        aTM = np.cross(aTE, kbar)
        aTM = aTM / np.linalg.norm(aTM)
    N = pw[0] * pw[1]
    delta = np.zeros(N, dtype=complex)
    delta[(N - 1) // 2] = 1
    pxy = te * aTE + tm * aTM
    return np.concatenate([delta * pxy[0], delta * pxy[1]])


def poynting_flux(g, c_out, kp, wl, eps_in, eps_out, only_total=True):
    """alternative.py:235-245."""
    k0 = TWO_PI / wl
    kzi = np.conj(csqrt(k0 ** 2 * eps_in - kp[0] ** 2 - kp[1] ** 2)) / k0
    sx, sy = np.split(c_out, 2)
    kx, ky, kz = k_vectors(g, kp, wl, eps_out)
    sz = -(kx * sx + ky * sy) / kz
    t = kz.real / np.real(kzi) * (np.abs(sx) ** 2 + np.abs(sy) ** 2 + np.abs(sz) ** 2)
    return np.sum(t) if only_total else (np.sum(t), t)


# --------------------------------------------------------------------------- #
# Crystal-level orchestration (khepri/crystal.py, khepri/layer.py)
# --------------------------------------------------------------------------- #
def make_structure(pw, layers, stack, lattice=None, epsi=1, epse=1):
    lattice = np.eye(2) if lattice is None else np.asarray(lattice, dtype=float)
    return {"pw": tuple(pw), "lattice": lattice, "epsi": epsi, "epse": epse,
            "layers": dict(layers), "stack": list(stack)}


def _structure_g(st):
    if "g" in st:                      # explicit (rotated / moire) basis
        return np.asarray(st["g"], dtype=float)
    return g_vectors(st["pw"], st["lattice"])


PATTERNED_BY_DOUBLING = None      # set to a slicing_pow to route patterned layers through layer_smatrix_doubling (tests only)


def solve_layer(spec, g, pw, kp, wl):
    """layer.py:145-194 -> dict(S, W, V, L, IC)."""
    Kx, Ky, _ = k_vectors(g, kp, wl)
    W0, V0 = free_space_modes(Kx, Ky)
    k0 = TWO_PI / wl
    kind = spec[0]
    if kind == "pixmap" and PATTERNED_BY_DOUBLING is not None:
        C = convolution_matrix(spec[1], pw)
        Pm, Qm = pq_matrices(Kx, Ky, C)
        return {"S": layer_smatrix_doubling(Pm, Qm, W0, V0, spec[2], k0, PATTERNED_BY_DOUBLING), "W": None, "V": None, "L": None,
                "IC": inv(C), "depth": spec[2]}
    if kind == "pixmap":
        C = convolution_matrix(spec[1], pw)          # recomputed per solve, as the reference does
        IC = inv(C)
        W, V, L = structured_layer_modes(Kx, Ky, C)
        S = layer_smatrix(W, V, W0, V0, L, spec[2], k0)
    elif kind == "analytical":                       # ("analytical", islands, eps_host, depth, lattice)
        C = analytical_convolution_matrix(spec[1], spec[3], pw, spec[4])
        IC = inv(C)
        W, V, L = structured_layer_modes(Kx, Ky, C)
        S = layer_smatrix(W, V, W0, V0, L, spec[2], k0)
    elif kind == "convmat":                          # precomputed C (analytic / test input)
        C = np.asarray(spec[1], dtype=complex)
        IC = inv(C)
        W, V, L = structured_layer_modes(Kx, Ky, C)
        S = layer_smatrix(W, V, W0, V0, L, spec[2], k0)
    elif kind == "uniform":
        W, V, L = uniform_layer_modes(Kx, Ky, spec[1])
        S = layer_smatrix(W, V, W0, V0, L, spec[2], k0)
        IC = 1 / spec[1]
    elif kind == "half_inc":
        S, W, V, L = halfspace_reflection(Kx, Ky, W0, V0, spec[1])
        IC = 1 / spec[1]
    elif kind == "half_trn":
        S, W, V, L = halfspace_transmission(Kx, Ky, W0, V0, spec[1])
        IC = 1 / spec[1]
    else:
        raise ValueError(kind)
    return {"S": S, "W": W, "V": V, "L": L, "IC": IC, "depth": 0.0 if kind.startswith("half") else spec[2]}


def stack_chain(n, layer_S, want_reverse=True):
    """layer.py:35-60 -> (prefix products, suffix products, Stot)."""
    Stot = identity_smatrix(n)
    prefix = []
    for S in layer_S:
        Stot = star(Stot, S)
        prefix.append(Stot.copy())
    suffix = None
    if want_reverse:
        Srev = identity_smatrix(n)
        suffix = []
        for S in reversed(layer_S[1:]):
            suffix.append(Srev.copy())
            Srev = star(S.copy(), Srev)
        suffix.append(Srev.copy())
        suffix.reverse()
    return prefix, suffix, Stot


def solve_structure(st, wl, kp, want_reverse=True):
    """crystal.py:131-206: half spaces are added, distinct layers solved once, chain built."""
    g = _structure_g(st)
    pw = st["pw"]
    specs = dict(st["layers"])
    specs["Sref"] = ("half_inc", st["epsi"])
    specs["Strans"] = ("half_trn", st["epse"])
    order = ["Sref", *st["stack"], "Strans"]
    solved = {name: solve_layer(specs[name], g, pw, kp, wl) for name in set(order)}
    n = 2 * g.shape[1]
    prefix, suffix, Stot = stack_chain(n, [solved[nm]["S"] for nm in order], want_reverse)
    depths = [solved[nm]["depth"] for nm in order]
    pos = list(np.cumsum(depths))
    pos[-1] = np.inf
    pos.insert(0, -np.inf)
    return {"order": order, "layers": solved, "prefix": prefix, "suffix": suffix,
            "Stot": Stot, "positions": pos, "g": g, "kp": kp, "wl": wl}


def source_kzi(st, wl, kp):
    k0 = TWO_PI / wl
    return np.conj(cmath.sqrt(k0 ** 2 * st["epsi"] - kp[0] ** 2 - kp[1] ** 2))


def flux_end(st, sol, te, tm, only_total=True):
    """crystal.py:363-396 -> (R, T)."""
    wl, kp, g = sol["wl"], sol["kp"], sol["g"]
    inc = incident_vector(st["pw"], te, tm, (kp[0], kp[1], source_kzi(st, wl, kp)))
    Wref = sol["layers"]["Sref"]["W"]
    Wtrn = sol["layers"]["Strans"]["W"]
    c1p = inv(Wref) @ inc
    T = poynting_flux(g, Wtrn @ sol["Stot"][1, 0] @ c1p, kp, wl, st["epsi"], st["epse"], only_total)
    R = poynting_flux(g, Wref @ sol["Stot"][0, 0] @ c1p, kp, wl, st["epsi"], st["epsi"], only_total)
    if only_total:
        return R.real, T.real
    return R, T


def solve_rt(st, wl, te=1.0, tm=1.0, theta=0.0, phi=0.0, kp=None):
    """One reference sweep iteration: set_source; solve; poynting_flux_end."""
    if kp is None:
        kp = kplanar(st["epsi"], wl, theta, phi)
    sol = solve_structure(st, wl, tuple(kp))
    return flux_end(st, sol, te, tm)


# --------------------------------------------------------------------------- #
# fields (khepri/fields.py, khepri/fourier.py:136-142, crystal.py:208-343)
# --------------------------------------------------------------------------- #
def eigenbasis(W, V):
    """fields.py:46-51."""
    return np.block([[W, W], [-V, V]])


def mode_amplitudes_in_gap(Sl, Sr, c1p):
    """fields.py:18-27."""
    n = len(c1p)
    cp = solve(np.eye(n) - Sl[1, 1] @ Sr[0, 0], Sl[1, 0] @ c1p)
    return cp, Sr[0, 0] @ cp


def fourier_fields(RI, LI, R0, amps, zbar):
    """fields.py:29-31, 53-62 (growing phasors clipped at 1e14)."""
    L = np.concatenate([np.exp(LI * zbar), np.exp(-LI * zbar)])
    big = np.abs(L) > 1e14
    L[big] /= np.abs(L[big]) / 1e14
    return np.split(RI @ (L * solve(RI, R0 @ np.concatenate(amps))), 4)


def longitudinal(sx, sy, ux, uy, Kx, Ky, IC):
    """fields.py:68-76."""
    uz = -1j * (Kx * sy - Ky * sx)
    rhs = Kx * uy - Ky * ux
    sz = -1j * (IC * rhs if np.isscalar(IC) else IC @ rhs)
    return sz, uz


def idft(s, kx, ky, x, y):
    """fourier.py:136-142."""
    ph = np.exp(1j * (kx[:, None] * x.ravel()[None, :] + ky[:, None] * y.ravel()[None, :]))
    return (s[:, None] * ph).sum(0).reshape(x.shape)


def locate(sol, z):
    """crystal.py:208-232."""
    pos = sol["positions"]
    i = int(np.searchsorted(pos, z) - 1)
    zr = z if z <= 0 else z - pos[i]
    return i, zr


def fields_fourier_at(st, sol, z, inc_eh):
    """crystal.py:234-283 -> (sx, sy, sz, ux, uy, uz) Fourier vectors."""
    i, zr = locate(sol, z)
    lay = sol["layers"][sol["order"][i]]
    wl, kp, g = sol["wl"], sol["kp"], sol["g"]
    Kx, Ky, _ = k_vectors(g, kp, wl)
    W0, V0 = free_space_modes(Kx, Ky)
    k0 = TWO_PI / wl
    ref = sol["layers"]["Sref"]
    c1p = np.split(solve(eigenbasis(ref["W"], ref["V"]), inc_eh), 2)[0]
    amps = mode_amplitudes_in_gap(sol["prefix"][i], sol["suffix"][i], c1p)
    sx, sy, ux, uy = fourier_fields(eigenbasis(lay["W"], lay["V"]), lay["L"], eigenbasis(W0, V0),
                                    amps, k0 * (lay["depth"] - zr))
    sz, uz = longitudinal(sx, sy, ux, uy, Kx, Ky, lay["IC"])
    return sx, sy, sz, ux, uy, uz


def fields_volume(st, sol, x, y, zs, te, tm, incident_fields=None):
    """crystal.py:285-343 -> (E, H), each (nz, 3, ny, nx).  incident_fields: explicit (E, H) Fourier vectors [4N]."""
    wl, kp, g = sol["wl"], sol["kp"], sol["g"]
    if incident_fields is None:
        e = incident_vector(st["pw"], te, tm, (kp[0], kp[1], source_kzi(st, wl, kp)))
        inc = np.concatenate([e, np.zeros_like(e)])
    else:
        inc = np.asarray(incident_fields, dtype=complex).reshape(-1)
    Kx, Ky, _ = k_vectors(g, kp, wl)
    k0 = TWO_PI / wl
    out = np.empty((len(zs), 6) + x.shape, dtype=complex)
    for iz, z in enumerate(zs):
        comps = fields_fourier_at(st, sol, z, inc)
        for c, s in enumerate(comps):
            out[iz, c] = idft(s, k0 * Kx, k0 * Ky, x, y)
    return out[:, :3], out[:, 3:]


def beam_amplitudes(fields, g, kp, x, y, bzs):
    """beams.py:164-191 (amplitudes_from_fields): samples (ny, nx, 2, 3) of a real-space source on the supercell points
    (x, y), divided by the Bloch phase of kp, transformed tile by tile with slow_dft (fourier.py:93-104) on the harmonics g,
    each tile normalised by n_tiles * NS (sic) and summed -> (4, N) = (Ex, Ey, Hx, Hy)_g."""
    fields = np.asarray(fields)
    ny, nx = fields.shape[:2]
    NS = ny // bzs[1]
    F = (fields / np.exp(1j * (kp[0] * x + kp[1] * y))[..., None, None])[..., :2].reshape(ny * nx, 4)
    ph = np.exp(-1j * (g[0][None, :] * x.reshape(-1, 1) + g[1][None, :] * y.reshape(-1, 1)))        # [pts, N]
    return (F.T @ ph) / (bzs[0] * bzs[1]) / NS


# --------------------------------------------------------------------------- #
# extended (twisted bilayer) RCWA  (khepri/extension.py, expansion.py:52-73)
# --------------------------------------------------------------------------- #
def rotation(theta):
    c, s = math.cos(theta), math.sin(theta)
    return np.array([[c, -s], [s, c]])


def minkowski_sum(g_lhs, g_rhs):
    """expansion.py:55-73: index = i_lhs * N + i_rhs."""
    return (g_lhs[:, :, None] + g_rhs[:, None, :]).reshape(2, -1)


def joint_block(blocks, kind):
    """extension.py:10-51 for one (n, n)=(2N, 2N) quadrant per shift.

    kind 0: joint harmonic = shift*N + r (block diagonal);  kind 1: r*N + shift.
    Vector halves (x, y) stay outermost.
    """
    N = blocks[0].shape[0] // 2
    out = np.zeros((2 * N * N, 2 * N * N), dtype=blocks[0].dtype)
    r = np.arange(N)
    for i, M in enumerate(blocks):
        j = i * N + r if kind == 0 else r * N + i
        for a in range(2):
            for b in range(2):
                out[np.ix_(a * N * N + j, b * N * N + j)] = M[a * N:(a + 1) * N, b * N:(b + 1) * N]
    return out


def joint_smatrix(S_list, kind):
    """extension.py:54-63."""
    return np.array([[joint_block([S[i, j] for S in S_list], kind) for j in range(2)] for i in range(2)])


def solve_extended_layer(spec, g_base, pw_base, g_other, kind, kp, wl):
    """extension.py:82-112: N shifted base solves scattered into the moire basis."""
    S_list = [solve_layer(spec, g_base, pw_base, (kp[0] + s[0], kp[1] + s[1]), wl)["S"] for s in g_other.T]
    return joint_smatrix(S_list, kind)


def solve_twisted(tw, wl, kp):
    """PRL_2021_BL notebook cells 4-8 via crystal.py:83-95,125-129.

    ``tw`` = {"pw", "g1", "g2" (rotated base g-vectors), "epsi", "epse",
              "layers": {name: (spec, which)} with which in {1, 2, None}, "stack"}.
    Layers with which=1 live on g1 (mode 1, shifts = g2); which=2 on g2 (mode 0,
    shifts = g1); which=None are plain layers on the moire basis.
    """
    g1, g2 = tw["g1"], tw["g2"]
    gm = minkowski_sum(g1, g2)
    n = 2 * gm.shape[1]
    pwm = (tw["pw"][0] ** 2, tw["pw"][1] ** 2)
    specs = dict(tw["layers"])
    specs["Sref"] = (("half_inc", tw["epsi"]), None)
    specs["Strans"] = (("half_trn", tw["epse"]), None)
    order = ["Sref", *tw["stack"], "Strans"]
    solved = {}
    for name in set(order):
        spec, which = specs[name]
        if which is None:
            solved[name] = solve_layer(spec, gm, pwm, kp, wl)
        elif which == 1:
            solved[name] = {"S": solve_extended_layer(spec, g1, tw["pw"], g2, 1, kp, wl)}
        else:
            solved[name] = {"S": solve_extended_layer(spec, g2, tw["pw"], g1, 0, kp, wl)}
    _, _, Stot = stack_chain(n, [solved[nm]["S"] for nm in order], want_reverse=False)
    return {"order": order, "layers": solved, "Stot": Stot, "g": gm, "kp": kp, "wl": wl}


# --------------------------------------------------------------------------- #
# tiny pixmap helpers for synthetic inputs (semantics of khepri/draw.py:19-56)
# --------------------------------------------------------------------------- #
def disc_pixmap(shape, eps_bg, center, radius, eps):
    xs = np.linspace(-0.5, 0.5, shape[0])[:, None]
    ys = np.linspace(-0.5, 0.5, shape[1])[None, :]
    pm = np.full(shape, float(eps_bg))
    pm[np.sqrt((xs - center[0]) ** 2 + (ys - center[1]) ** 2) < radius] = eps
    return pm


def rect_pixmap(shape, eps_bg, center, wh, eps, base=None):
    xs = np.linspace(-0.5, 0.5, shape[0])[:, None]
    ys = np.linspace(-0.5, 0.5, shape[1])[None, :]
    pm = np.full(shape, float(eps_bg)) if base is None else base
    x0, y0 = center[0] - wh[0] / 2, center[1] - wh[1] / 2
    m = (xs >= x0) & (xs <= x0 + wh[0]) & (ys >= y0) & (ys <= y0 + wh[1])
    pm[m] = eps
    return pm


# --------------------------------------------------------------------------- #
# band post-processing (khepri/eigentricks.py)
# --------------------------------------------------------------------------- #
def scattering_splitlr(S):
    """Pencil (Sl, Sr) of a flat 2n x 2n S-matrix.  eigentricks.py:5-21."""
    h = S.shape[0] // 2
    I = np.eye(h, dtype=S.dtype)
    Z = np.zeros((h, h), dtype=S.dtype)
    Sl = np.block([[S[:h, :h], Z], [S[h:, :h], -I]])
    Sr = np.block([[I, -S[:h, h:]], [Z, -S[h:, h:]]])
    return Sl, Sr


def scattering_eigenvalues(S):
    """Generalized eigenpairs of (Sl, Sr) with scipy's QZ, as the reference.  eigentricks.py:29-40."""
    from scipy.linalg import eig as geig
    Sl, Sr = scattering_splitlr(S)
    if np.any(np.isnan(Sr)) or np.any(np.isnan(Sl)):
        return None
    return geig(Sl, Sr)


def scattering_det(S):
    """eigentricks.py:23-27."""
    h = S.shape[0] // 2
    return np.linalg.det(S[:h, :h] - S[:h, h:] @ inv(S[h:, h:]) @ S[h:, :h]) * np.linalg.det(S[h:, h:])


def flat_smatrix(S4):
    """(2,2,n,n) -> (2n,2n) block matrix."""
    return np.block([[S4[0, 0], S4[0, 1]], [S4[1, 0], S4[1, 1]]])
