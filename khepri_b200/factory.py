"""Device builders of the reference (khepri/factory.py:3-24): the woodpile of BASELINE configs[2].

Input generation on the host: four pixmap layers A, B, C, D (a rod, the rod transposed, the shifted pair of rods, its
transpose) handed to a Crystal whose solves run on the GPU path like any other pixmap stack
(examples/crystal_api/woodpile.py:36-39, 85 doubles the cell with ``cl.Stot = redheffer_product(cl.Stot, cl.Stot)``).
"""
from .crystal import Crystal
from .draw import Drawing


def make_woodpile(rods_w, rods_eps, rods_shift, rods_height, pw, resolution=(256, 256), engine=None):
    """Same arguments and layer names as the reference; ``engine`` optionally selects the Engine (default: shared one)."""
    rod = Drawing(resolution, 1)
    rod.rectangle((0, 0), (1, rods_w), rods_eps)
    pair = Drawing(resolution, 1)
    pair.rectangle((0, rods_shift), (1, rods_w), rods_eps)
    pair.rectangle((0, -rods_shift), (1, rods_w), rods_eps)

    cl = Crystal(pw, engine=engine)
    for name, canvas in (("A", rod.canvas()), ("B", rod.canvas().T), ("C", pair.canvas()), ("D", pair.canvas().T)):
        cl.add_layer_pixmap(name, canvas, rods_height)
    stack = ["A", "B", "C", "D"]
    cl.set_device(stack, [False] * len(stack))
    return cl
