"""Minimal pixmap generator with khepri.draw.Drawing's semantics for discs and rectangles
(arrays indexed a[x, y], unit cell [-0.5, 0.5]^2).  Input synthesis only -- it runs on the host."""
import numpy as np


class Drawing:
    def __init__(self, shape, epsilon_background, lattice=None):
        self.geometric_description = []
        self.lattice = np.eye(2) if lattice is None else np.asarray(lattice, dtype=float)
        self.background = epsilon_background
        self._canvas = np.ones(shape) * epsilon_background
        nX, nY = np.meshgrid(np.linspace(-0.5, 0.5, shape[0]), np.linspace(-0.5, 0.5, shape[1]), indexing="ij")
        self.X = self.lattice[0, 0] * nX + self.lattice[1, 0] * nY
        self.Y = self.lattice[0, 1] * nX + self.lattice[1, 1] * nY

    def disc(self, xy, radius, epsilon):
        self._canvas[np.sqrt((self.X - xy[0]) ** 2 + (self.Y - xy[1]) ** 2) < radius] = epsilon
        self.geometric_description.append({"type": "disc", "params": [0.5 + xy[0], 0.5 + xy[1], radius], "epsilon": epsilon})

    def rectangle(self, xy, wh, epsilon):
        x, y = xy[0] - wh[0] / 2, xy[1] - wh[1] / 2
        inside = (self.X >= x) & (self.X <= x + wh[0]) & (self.Y >= y) & (self.Y <= y + wh[1])
        self._canvas[inside] = epsilon
        self.geometric_description.append({"type": "rectangle", "params": [0.5 + x, 0.5 + y, 0.5 + x + wh[0], 0.5 + y + wh[1]], "epsilon": epsilon})

    def islands(self):
        """Geometric description for Crystal.add_layer_analytical (draw.py:108-109)."""
        return self.geometric_description

    def canvas(self):
        return self._canvas.copy()
