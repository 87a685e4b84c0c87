"""Small host helpers with the reference's names (khepri/tools.py)."""
from cmath import sqrt as _csqrt
from math import cos, sin

import numpy as np

twopi = 2 * np.pi


def compute_kplanar(eps_inc, wavelength, theta_deg=0.0, phi_deg=0.0):
    """In-plane wavevector of the source; angles in DEGREES (tools.py:17-21)."""
    phi, theta = np.deg2rad(phi_deg), np.deg2rad(theta_deg)
    kp = np.array([np.cos(phi), np.sin(phi)], dtype=complex)
    return kp * _csqrt(eps_inc) * 2 * np.pi / wavelength * np.sin(theta)


def rotation_matrix(theta):
    return np.array([[cos(theta), -sin(theta)], [sin(theta), cos(theta)]])


def unitcellarea(a1, a2):
    return abs(a1[0] * a2[1] - a1[1] * a2[0])


def reciproc(a1, a2):
    """Reciprocal basis (tools.py:64-69)."""
    coef = twopi / (a1[0] * a2[1] - a1[1] * a2[0])
    return (a2[1] * coef, -a2[0] * coef), (-a1[1] * coef, a1[0] * coef)


def block2dense(block_matrix):
    """(2,2,n,n) block S-matrix -> (2n,2n) (tools.py:26-30)."""
    b = np.asarray(block_matrix)
    return np.swapaxes(b, 1, 2).reshape(b.shape[0] * b.shape[2], b.shape[1] * b.shape[3])


def convolution_matrix(structure, harmonics, engine=None):
    """tools.py:33-35 on the GPU: pruned DFT + Toeplitz gather.  Returns a numpy array."""
    from .engine import Engine
    eng = engine or Engine.default()
    return eng.convmat(np.asarray(structure), harmonics)[0].cpu().numpy()


def convolution_matrix_fourier(fourier_coefficients, harmonics, engine=None):
    """tools.py:38-56 on the GPU (bit-exact index gather)."""
    from .engine import Engine
    eng = engine or Engine.default()
    return eng.toeplitz_gather(np.asarray(fourier_coefficients, dtype=np.complex128), harmonics).cpu().numpy()
