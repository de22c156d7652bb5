"""Post-PDE membrane updates on device-resident vectors.

``utils.update_pde_variables`` (src/knpemi/utils.py:238-293) does three things per subdomain
after every PDE solve; with the PDE coefficient vectors in HBM each is one streaming kernel of
libknpemi_b200 instead of a host round trip:

* eliminated-ion concentration (utils.py:247-267)
      c_elim = -(1/z_e) * rho_z * rho_tag + sum_k -(1/z_e) * z_k * c_k
  -> :func:`eliminated_ion_terms` gives the constant and the coefficients in the reference's
  order of summation, :func:`affine_combine` evaluates them over a bulk vector, and
  ``MembraneModel.set_from_device_affine`` takes the membrane trace of the same combination
  straight into a parameter column (what ``update_ode_variables`` pushes for the last ion,
  utils.py:219-228);
* membrane potential ``phi_M = tr(phi_i) - tr(phi_e)`` (utils.py:288-291)
  -> ``MembraneModel.set_membrane_potential_from_device``;
* Nernst potentials (utils.py:271-281) are UFL expressions of the weak forms: PDE side.
"""
from __future__ import annotations

import ctypes as C

from . import _cabi
from ._cabi import KemError, check

MAX_TERMS = 8


def eliminated_ion_terms(ion_list, rho_z, rho_tag):
    """(a0, [coef_k]) of ``c_elim = a0 + sum_k coef_k * c_k`` for the ions ``ion_list[:-1]``, the
    last entry of ``ion_list`` being the eliminated ion -- formed like utils.py:249,258:
    ``a0 = -(1.0 / z_e) * rho_z * rho_tag``, ``coef_k = -(1.0 / z_e) * z_k``."""
    z_e = float(ion_list[-1]['z'])
    a0 = -(1.0 / z_e) * float(rho_z) * float(rho_tag)
    coefs = [-(1.0 / z_e) * float(ion['z']) for ion in ion_list[:-1]]
    return a0, coefs


def _pack(terms):
    if len(terms) > MAX_TERMS:
        raise KemError(f"at most {MAX_TERMS} terms")
    coef = (C.c_double * max(len(terms), 1))(*[float(c) for c, _ in terms])
    ptrs = (C.c_void_p * max(len(terms), 1))(*[int(p) for _, p in terms])
    return coef, ptrs


def affine_combine(dev, n, out_ptr, a0, terms):
    """out[i] = a0 + sum_k coef_k * in_k[i] over n DOFs on device `dev`; `terms` is a list of
    (coefficient, device pointer); `out_ptr` a device pointer.  Returns when the result is there."""
    coef, ptrs = _pack(terms)
    check(_cabi.lib().kem_device_affine_combine(int(dev), int(n), C.c_void_p(int(out_ptr)), float(a0),
                                                len(terms), coef, ptrs), "kem_device_affine_combine")
