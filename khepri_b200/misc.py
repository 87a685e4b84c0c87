"""khepri/misc.py helpers used by the field-map sweeps."""
import numpy as np


def coords(xmin, xmax, ymin, ymax, zmin, zmax, resolution):
    """misc.py:28-32: 'xy' meshgrid + z vector."""
    x = np.linspace(xmin, xmax, resolution[0])
    y = np.linspace(ymin, ymax, resolution[1])
    z = np.linspace(zmin, zmax, resolution[2])
    return *np.meshgrid(x, y, indexing="xy"), z


def poynting_vector(E, H, axis=-1):
    return np.cross(E, np.conj(H), axis=axis)


def ensure_array(x):
    return np.array(x)
