"""``knpemi.odeSolver`` of the reference (src/knpemi/odeSolver.py), B200 backend."""
from knpemi_b200.odeSolver import KemError, MembraneModel, NonFiniteStateError  # noqa: F401

__all__ = ["MembraneModel"]
