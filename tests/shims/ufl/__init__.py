"""Stub of ``ufl`` -- TEST INFRASTRUCTURE ONLY (``from ufl import ln``, src/knpemi/utils.py:11-13;
used by the PDE-side update only)."""
import math as _math


def ln(x):
    return _math.log(x)
