"""Minimal stand-ins for the two dolfinx objects MembraneModel touches.

The reference needs only ``Q.tabulate_dof_coordinates()`` (odeSolver.py:32) from
the function space and ``u.x.array`` (odeSolver.py:142,159,164) from a Function.
dolfinx is not required by this backend; these duck types let the stage run
stand-alone (benchmarks, tests, the calibration driver).
"""
from __future__ import annotations

import numpy as np


class PointSpace:
    """Function-space stand-in: a list of DOF coordinates ``float64[N, 3]``."""

    def __init__(self, coordinates):
        self._x = np.ascontiguousarray(coordinates, dtype=np.float64)
        if self._x.ndim != 2:
            raise ValueError("coordinates must be [N, gdim]")

    def tabulate_dof_coordinates(self):
        return self._x


class _Vector:
    def __init__(self, array):
        self.array = array


class ArrayFunction:
    """Function stand-in: ``u.x.array`` is a writable 1-D float64 array."""

    def __init__(self, n_or_array, function_space=None):
        if isinstance(n_or_array, (int, np.integer)):
            a = np.zeros(int(n_or_array), dtype=np.float64)
        else:
            a = np.asarray(n_or_array, dtype=np.float64)
        self.x = _Vector(a)
        self.function_space = function_space
