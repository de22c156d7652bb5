"""Stub of ``dolfinx`` -- TEST INFRASTRUCTURE ONLY.

dolfinx is not installable in this image.  The membrane side of the
reference touches very little of it: ``dolfinx.common.Timer`` (src/knpemi/odeSolver.py:104),
``dolfinx.fem.Function(Q)`` with ``.x.array`` / ``.x.scatter_forward()`` / ``.function_space``
(src/knpemi/utils.py:136,190-191,217), ``float(dolfinx.fem.Constant)`` (utils.py:124-129) and
the name ``dolfinx.mesh.MeshTags`` in a type annotation (utils.py:18).  This stub supplies
exactly that, so that the reference's own ``odeSolver.py`` and ``utils.py`` can be imported from
/root/reference and executed verbatim (tests/golden/make_glue_transcript.py,
tests/test_reference_glue.py).  Nothing under knp-emi-fenics-x_b200/ imports it.
"""
import time as _time
import types as _types

import numpy as _np


class _Timer:
    def __init__(self, name=""):
        self.name, self._t0 = name, None

    def start(self):
        self._t0 = _time.perf_counter()

    def stop(self):
        return _time.perf_counter() - (self._t0 or _time.perf_counter())


class _Vector:
    def __init__(self, n):
        self.array = _np.zeros(n, dtype=_np.float64)

    def scatter_forward(self):
        pass


class _Function:
    def __init__(self, V, name=None):
        self.function_space = V
        self.name = name or "f"
        self.x = _Vector(len(V.tabulate_dof_coordinates()))


class _Constant:
    def __init__(self, domain, value):
        self.value = value

    def __float__(self):
        return float(self.value)


class _MeshTags:
    pass


common = _types.SimpleNamespace(Timer=_Timer)
fem = _types.SimpleNamespace(Function=_Function, Constant=_Constant, Expression=None)
mesh = _types.SimpleNamespace(MeshTags=_MeshTags)
