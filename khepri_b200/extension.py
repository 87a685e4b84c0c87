"""Extended (twisted bilayer) RCWA descriptor.  Mirrors khepri/extension.py:66-80; the shifted base
solves and the joint-subspace scatter (extension.py:82-112, 10-63) run on the GPU."""
from .layer import Layer


class ExtendedLayer:
    def __init__(self, expansion, base):
        if getattr(expansion, "expansion_lhs", None) is base.expansion:
            self.mode = 1
            self.gs = expansion.expansion_rhs.g_vectors
        elif getattr(expansion, "expansion_rhs", None) is base.expansion:
            self.mode = 0
            self.gs = expansion.expansion_lhs.g_vectors
        else:
            raise NotImplementedError("Base layer expansion should be in the extented expansion.")
        if not isinstance(base, Layer):
            raise NotImplementedError("ExtendedLayer over a whole Crystal is not supported; wrap Layer objects.")
        self.expansion = expansion
        self.base = base
        self.depth = base.depth
        self.fields = base.fields
        self.IC = 1.0
