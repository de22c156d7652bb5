"""Seeded synthetic workloads shared by the tests and bench.py.

Values the reference run scripts put into the uniform (U) and per-DOF PDE input
(P) parameter slots:
  hh_ideal      examples/idealized_geometries/run_2D.py:174-195,237-249,263-266
  hh_tissue /   examples/local_astrocyte_depolarization/run_stim_duration.py:216-242
  glial_*       (tissue units: Cm = 1, psi = F/(R T) with F=96500e3, R=8.315e3, T=307e3)
  calibration / hh_test: module defaults are directly usable.
"""
from __future__ import annotations

import numpy as np

_PSI_TISSUE = 96500e3 / (8.315e3 * 307e3)
_GLIAL = dict(Cm=1.0, psi=_PSI_TISSUE, z_Na=1.0, z_K=1.0, z_Cl=-1.0,
              Na_i=15.775818906083778, Na_e=144.60625137617149, K_i=99.3100014897692,
              K_e=3.092970607490389, Cl_e=133.62525154406637, Cl_i=5.203660274163705)

SETUP = {
    "hh_ideal": dict(dt=1.0e-4, stim=10.0, uniform=dict(
        Cm=0.02, psi=96485.0 / (8.314 * 300.0), z_Na=1.0, z_K=1.0, z_Cl=-1.0), varying=dict(
        Na_i=12.838513108648856, Na_e=100.71925900027354, K_i=124.15397583491901,
        K_e=3.3236967382705265, Cl_e=100.71925900027354 + 3.3236967382705265,
        Cl_i=12.838513108648856 + 124.15397583491901)),
    "hh_tissue": dict(dt=0.1, stim=5.0, uniform=dict(
        Cm=1.0, psi=_PSI_TISSUE, z_Na=1.0, z_K=1.0, z_Cl=-1.0), varying=dict(
        Na_i=12.838513108648856, Na_e=100.71925900027354, K_i=124.15397583491901,
        K_e=3.3236967382705265, Cl_e=104.04295573854407, Cl_i=136.99248894356787)),
    "glial_tissue": dict(dt=0.1, stim=0.0,
                         uniform={k: _GLIAL[k] for k in ("Cm", "psi", "z_Na", "z_K", "z_Cl")},
                         varying={k: _GLIAL[k] for k in ("Na_i", "Na_e", "K_i", "K_e", "Cl_e", "Cl_i")}),
    "glial_bench": dict(dt=0.1, stim=0.0,
                        uniform={k: _GLIAL[k] for k in ("Cm", "psi", "z_Na", "z_K", "z_Cl")},
                        varying={k: _GLIAL[k] for k in ("Na_i", "Na_e", "K_i", "K_e", "Cl_e", "Cl_i")}),
    "calibration": dict(dt=0.1, stim=2.0, uniform={}, varying={}),
    "hh_test": dict(dt=0.1, stim=0.5, uniform={}, varying={}),
}


def builtin(name):
    from knpemi_b200.models import BUILTIN
    return BUILTIN[name]


def synthetic_tables(name, n, seed=20240611):
    """AoS tables + coordinates + stimulus mask of the synthetic workload.

    States: defaults tiled, V-like states perturbed by 5 %, gates clipped-perturbed
    by 0.02; varying parameters perturbed by 2 %; coordinates uniform in
    (0, 62e-6)^3; stimulus mask x[0] < 20e-6 (about 32 % of the DOFs).
    Returns (states[n,ns], params[n,np], X[n,3], mask[n]); the stimulus amplitude
    column is left at the module default (the step call applies it)."""
    ode = builtin(name)
    rng = np.random.default_rng(seed)
    y0, p0 = ode.init_state_values(), ode.init_parameter_values()
    cfg = SETUP[name]
    for k, v in cfg["uniform"].items():
        p0[ode.parameter_indices(k)] = v
    ns = len(y0)
    states = np.tile(y0, (n, 1))
    states *= 1.0 + 0.05 * rng.uniform(-1, 1, (n, ns))
    if ns >= 4:
        states[:, :3] = np.clip(np.tile(y0[:3], (n, 1)) + 0.02 * rng.uniform(-1, 1, (n, 3)),
                                1e-6, 1 - 1e-6)
    params = np.tile(p0, (n, 1))
    for k, v in cfg["varying"].items():
        params[:, ode.parameter_indices(k)] = v * (1.0 + 0.02 * rng.uniform(-1, 1, n))
    X = rng.uniform(0.0, 62e-6, (n, 3))
    mask = X[:, 0] < 20e-6
    return np.ascontiguousarray(states), np.ascontiguousarray(params), X, mask


def load_tables(model, states, params):
    """Push AoS host tables into a MembraneModel (uniform columns stay uniform)."""
    for c in range(states.shape[1]):
        model._set_column(0, c, np.ascontiguousarray(states[:, c]))
    for c in range(params.shape[1]):
        col = params[:, c]
        if np.all(col == col[0]):
            from knpemi_b200._cabi import check
            check(model._lib.kem_set_uniform(model._h, 1, c, float(col[0])), "kem_set_uniform")
        else:
            model._set_column(1, c, np.ascontiguousarray(col))
