"""Scheme O1 (RK4 x 25, the product's scheme) against scheme O2 (LSODA at the
reference's tolerances rtol 1e-8 / atol 1e-10, cold start per row: odeSolver.py:116-120)
and against a tight solution (RK4 x 800).

O2 bounds, it does not pin: scipy's LSODA is the same algorithm family as the
reference's numbalsoda, different code.  Measured in this container after three
stimulated PDE steps (relative to max(|y|, 1e-3 column max)):

    model         RK4x25 error   LSODA(1e-8) error
    hh_tissue       1.3e-7          3.5e-8
    hh_ideal        3.8e-9          3.2e-8
    glial_tissue    2.1e-13         2.9e-8
    calibration     2.4e-9          3.7e-8

so the fixed-step scheme sits at the accuracy level of the integrator it replaces."""
import numpy as np
import pytest

from ducks_for_tests import Space
from workloads import SETUP, builtin, synthetic_tables

BOUND = 5e-7


@pytest.mark.parametrize("name", ["hh_tissue", "hh_ideal", "glial_tissue", "calibration"])
def test_rk4_at_lsoda_accuracy_level(name):
    from oracle.membrane_oracle import OracleMembraneModel
    ode = builtin(name)
    S, P, X, mask = synthetic_tables(name, 5, seed=3)
    rk4, tight, lsoda = (OracleMembraneModel(ode, None, 1, Space(X), oracle_name=name, n_sub=k)
                         for k in (25, 800, 25))
    for m in (rk4, tight, lsoda):
        m.states[:] = S
        m.parameters[:] = P
    stim = {"stim_amplitude": SETUP[name]["stim"]}
    loc = lambda x: x[0] < 20e-6       # noqa: E731
    dt = SETUP[name]["dt"]
    for _ in range(3):
        rk4.step_lsoda(dt, stim, loc)
        tight.step_lsoda(dt, stim, loc)
        lsoda.step_lsoda_scipy(dt, stim, loc)
    ref = tight.states
    scale = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max(axis=0))
    err_rk4 = (np.abs(rk4.states - ref) / scale).max()
    err_lsoda = (np.abs(lsoda.states - ref) / scale).max()
    assert err_rk4 < BOUND and err_lsoda < BOUND, (err_rk4, err_lsoda)
    assert (np.abs(rk4.states - lsoda.states) / scale).max() < BOUND
    assert rk4.time == pytest.approx(lsoda.time)
