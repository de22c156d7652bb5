"""Import-path shim: makes ``from knpemi.odeSolver import MembraneModel`` (reference
src/knpemi/__init__.py:1, utils.py:9, run_calibration.py:8) resolve to the B200 backend.

Two ways to use it:

* stand-alone (no reference checkout, e.g. the calibration driver): put this directory's
  parent (``knp-emi-fenics-x_b200/compat``) on ``sys.path``;
* under the full reference: keep the reference's ``knpemi`` package and replace only its
  ``odeSolver.py`` by the one-line re-export shown in INTEGRATION.md (the reference's
  ``__init__`` imports dolfinx-dependent modules this shim does not provide).
"""
from .odeSolver import MembraneModel  # noqa: F401

__all__ = ["MembraneModel"]
