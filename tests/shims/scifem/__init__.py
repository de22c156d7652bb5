"""Stub of ``scifem`` -- TEST INFRASTRUCTURE ONLY (see tests/shims/dolfinx).  The reference calls
``scifem.interpolation.interpolate_to_surface_submesh`` inside ``interpolate_to_membrane``
(src/knpemi/utils.py:194-203), which the glue tests replace by synthetic traces."""
import types as _types


def _unavailable(*a, **k):
    raise NotImplementedError("scifem is stubbed: the trace interpolation is outside the membrane-ODE path")


interpolation = _types.SimpleNamespace(interpolate_to_surface_submesh=_unavailable)
