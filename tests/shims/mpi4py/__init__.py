"""Stub of ``mpi4py`` -- TEST INFRASTRUCTURE ONLY (``from mpi4py import MPI``, src/knpemi/utils.py:5)."""
import types as _types

MPI = _types.SimpleNamespace(COMM_WORLD=_types.SimpleNamespace(rank=0, size=1))
