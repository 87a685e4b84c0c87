"""khepri_b200 -- B200-native batched RCWA solve engine behind the khepri.crystal.Crystal API.

Only the hot path named in SURVEY.md §8 lives here: the Crystal / Layer / Expansion host mirror and
the sm_100a kernels (csrc/) reached through the C ABI of include/khepri_b200.h.  Importing the
package does not need a GPU; constructing an Engine (or solving a Crystal) does, and fails loudly
without one -- there is no CPU fallback.
"""
from .crystal import Crystal, Multilayer
from . import beams
from . import eigentricks
from . import factory
from .draw import Drawing
from .engine import Engine
from .expansion import Expansion
from .extension import ExtendedLayer
from .layer import Field, Formulation, Layer
from ._lib import KhepriError

__all__ = ["beams", "eigentricks", "Crystal", "Multilayer", "Drawing", "Engine", "Expansion", "ExtendedLayer", "Field", "Formulation", "Layer", "KhepriError"]
