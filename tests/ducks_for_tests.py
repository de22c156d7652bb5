"""Duck-typed dolfinx stand-ins for the tests (independent of the product's ducks.py)."""
import numpy as np


class Space:
    def __init__(self, X):
        self.X = np.ascontiguousarray(X, dtype=np.float64)

    def tabulate_dof_coordinates(self):
        return self.X


class _Vec:
    def __init__(self, a):
        self.array = a


class Func:
    def __init__(self, a):
        self.x = _Vec(np.array(a, dtype=np.float64))


class FloatLike:
    """Stands in for dolfinx.fem.Constant: only float() works (utils.py:124-129)."""

    def __init__(self, v):
        self.v = v

    def __float__(self):
        return float(self.v)
