"""Harmonic basis (host side, numpy).  Mirrors khepri/expansion.py: same attributes and conventions."""
import numpy as np

from .tools import reciproc, rotation_matrix, unitcellarea


def generate_expansion_indices(pw):
    """(2, N) integer harmonics, x index fastest (expansion.py:4-16)."""
    assert pw[0] % 2 == 1
    assert pw[1] % 2 == 1
    p = np.arange(pw[0]) - (pw[0] - 1) // 2
    q = np.arange(pw[1]) - (pw[1] - 1) // 2
    return np.stack([np.tile(p, pw[1]), np.repeat(q, pw[0])])


def kz_from_kplanar(kx, ky, k0, epsilon):
    """expansion.py:19-26."""
    arg = k0 ** 2 * np.conj(epsilon) - kx ** 2 - ky ** 2
    return np.conj(np.sqrt(arg.astype("complex")))


class Expansion:
    def __init__(self, pw, lattice=None):
        self.pw = pw
        if lattice is None:
            lattice = np.asarray([[1, 0], [0, 1]])
        lattice = np.asarray(lattice)
        self.a = np.linalg.norm(lattice[0])
        self.reciprocal = np.asarray(reciproc(lattice[0], lattice[1]))
        self.expansion_indices = generate_expansion_indices(pw)
        m, n = self.expansion_indices
        self._g_vectors = self.reciprocal[0][:, None] * m[None, :] + self.reciprocal[1][:, None] * n[None, :]
        self.sigma = unitcellarea(*lattice)

    @property
    def g_vectors(self):
        return self._g_vectors.copy()

    def g_vectors_expanded(self, mul):
        """g-vectors of a mul-times larger harmonic grid (expansion.py:86-92)."""
        epw = [e * mul if e > 1 else 1 for e in self.pw]
        m, n = generate_expansion_indices(epw)
        g = self.reciprocal[0][:, None] * m[None, :] + self.reciprocal[1][:, None] * n[None, :]
        return g[0], g[1], epw

    def k_vectors(self, k_parallel, wavelength, epsilon=1):
        """Normalised (Kx, Ky, Kz), shape (3, N) complex (expansion.py:43-50, 95-97)."""
        k0 = 2 * np.pi / wavelength
        kv = np.zeros((3, int(np.prod(self.pw))), dtype=np.complex128)
        kv[0:2, :] = np.asarray(k_parallel)[:, None] + self._g_vectors
        kv[2, :] = kz_from_kplanar(kv[0], kv[1], k0, epsilon)
        kv /= k0
        self._k_vectors = kv
        return kv

    def rotate(self, angle_rad):
        self._g_vectors = rotation_matrix(angle_rad) @ self._g_vectors

    def __add__(self, rhs):
        """Minkowski sum; joint index = i_self * N_rhs + i_rhs (expansion.py:55-73)."""
        g_sum = (self._g_vectors[:, :, None] + rhs._g_vectors[:, None, :]).reshape(2, -1)
        e = Expansion((self.pw[0] ** 2, self.pw[1] ** 2))
        e._g_vectors = g_sum
        e._base_g_vectors = np.stack((self.g_vectors, rhs.g_vectors))
        e.expansion_lhs = self
        e.expansion_rhs = rhs
        return e
