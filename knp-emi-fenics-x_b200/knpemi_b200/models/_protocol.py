"""Model-plugin protocol helpers.

A membrane model module, as consumed by ``MembraneModel`` (reference
src/knpemi/odeSolver.py:8-49), exports

* ``init_state_values(**overrides)``      -> float64[ns]
* ``init_parameter_values(**overrides)``  -> float64[np]
* ``state_indices(*names)`` / ``parameter_indices(*names)`` -> int | list[int]
* ``rhs_numba``  (C signature ``void(double t, double* y, double* dy, double* p)``,
  the ``lsoda_sig`` of e.g. examples/idealized_geometries/mm_hh.py:133,139)

The reference modules are Gotran-generated and repeat the name->index tables in
every function (mm_hh.py:7-131).  Here the four metadata functions are built
once from two ordered ``(name, default)`` tables; behaviour (return types,
``ValueError`` on unknown names) follows mm_hh.py:23-31,80-88,96-104,123-131.
"""
from __future__ import annotations

import numpy as np


def table_functions(states, parameters):
    """Return the four metadata functions of the plugin protocol."""
    s_names = [n for n, _ in states]
    p_names = [n for n, _ in parameters]
    s_ind = {n: i for i, n in enumerate(s_names)}
    p_ind = {n: i for i, n in enumerate(p_names)}
    s_def = np.array([v for _, v in states], dtype=np.float64)
    p_def = np.array([v for _, v in parameters], dtype=np.float64)

    def _init(defaults, index, what):
        def init(**values):
            out = defaults.copy()
            for name, value in values.items():
                if name not in index:
                    raise ValueError("{0} is not a {1}.".format(name, what))
                out[index[name]] = value
            return out
        return init

    def _indices(index, what):
        def indices(*names):
            found = []
            for name in names:
                if name not in index:
                    raise ValueError("Unknown {0}: '{1}'".format(what, name))
                found.append(index[name])
            return found if len(found) > 1 else found[0]
        return indices

    init_state_values = _init(s_def, s_ind, "state")
    init_state_values.__doc__ = "Default state vector; keyword overrides by name."
    init_parameter_values = _init(p_def, p_ind, "parameter")
    init_parameter_values.__doc__ = "Default parameter vector; keyword overrides by name."
    return (init_state_values, init_parameter_values,
            _indices(s_ind, "state"), _indices(p_ind, "param"))


class rhs_cfunc:
    """Lazy numba ``cfunc`` wrapper for a model right-hand side.

    Gives a builtin model the same ``rhs_numba.address`` / ``.ctypes`` surface
    the reference modules have (mm_hh.py:138-139) without paying the LLVM
    compile at import time and without requiring ``numbalsoda``.  The CUDA
    backend never calls this object: it reads the Python *source* of the
    decorated function (``knpemi_b200.codegen``).
    """

    def __init__(self, pyfunc):
        self._pyfunc = pyfunc
        self.__name__ = pyfunc.__name__
        self.__doc__ = pyfunc.__doc__
        self._compiled = None

    def _get(self):
        if self._compiled is None:
            from numba import cfunc, types
            sig = types.void(types.double, types.CPointer(types.double),
                             types.CPointer(types.double), types.CPointer(types.double))
            self._compiled = cfunc(sig, nopython=True)(self._pyfunc)
        return self._compiled

    @property
    def address(self):
        return self._get().address

    @property
    def ctypes(self):
        return self._get().ctypes

    def __call__(self, t, states, values, parameters):
        return self._pyfunc(t, states, values, parameters)
