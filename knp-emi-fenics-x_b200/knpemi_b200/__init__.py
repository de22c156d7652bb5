"""knpemi_b200 -- B200-native membrane-ODE stage behind the reference's MembraneModel API.

``from knpemi_b200.odeSolver import MembraneModel`` is the drop-in for
``from knpemi.odeSolver import MembraneModel`` (reference src/knpemi/__init__.py:1).
"""
__version__ = "0.1.0"
