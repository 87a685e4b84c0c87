"""Reference-named helpers of khepri/alternative.py that callers use directly."""
import logging
from math import prod

import numpy as np


def incident(pw, te_pol, tm_pol, k_vector, normalize=True):
    """2N source vector: delta at the central harmonic times the polarisation (alternative.py:101-128)."""
    logging.debug(f"Building vector with plane wave {pw=}, {te_pol=}, {tm_pol=}, {k_vector=}")
    if normalize:
        nrm = np.linalg.norm((abs(te_pol), abs(tm_pol)))
        te_pol, tm_pol = te_pol / nrm, tm_pol / nrm
    kvec = np.asarray(k_vector, dtype=np.complex128)
    kbar = kvec / np.linalg.norm(kvec)
    if abs(np.linalg.norm(kvec[:2])) < 1e-8:
        aTE, aTM = np.array([1, 0, 0]), np.array([0, 1, 0])
    else:
        aTE = -np.cross(np.array([0, 0, -1], dtype=np.complex128), kbar)
        aTE = aTE / np.linalg.norm(aTE)
        aTM = np.cross(aTE, kbar)
        aTM = aTM / np.linalg.norm(aTM)
    N = prod(pw)
    delta = np.zeros(N, dtype=np.complex128)
    delta[(N - 1) // 2] = 1
    pxy = te_pol * aTE + tm_pol * aTM
    return np.hstack([delta * pxy[0], delta * pxy[1]])


def scattering_identity(pw, block=False):
    """alternative.py:220-232."""
    n = 2 * prod(pw)
    eye, zero = np.eye(n), np.zeros((n, n))
    if block:
        return np.asarray([[zero, eye], [eye, zero]]).astype("complex")
    return np.block([[zero, eye], [eye, zero]]).astype("complex")


def redheffer_product(SA, SB, engine=None):
    """Star product of two (2,2,n,n) S-matrices on the GPU (alternative.py:19-30)."""
    from .engine import Engine
    eng = engine or Engine.default()
    return eng.star(np.asarray(SA, dtype=np.complex128), np.asarray(SB, dtype=np.complex128)).cpu().numpy()
