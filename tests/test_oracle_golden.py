"""The CPU oracle against the committed golden vectors (made from the reference's
own numba cfuncs by tests/golden/make_golden.py).  Bit-exact: same IEEE operations,
same glibc libm."""
import os

import numpy as np
import pytest

from conftest import MODEL_NAMES
from oracle import cpu_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_oracle_rhs_matches_reference_points(name):
    g = np.load(os.path.join(GOLDEN, f"rhs_{name}.npz"))
    for k in range(len(g["t"])):
        dy, p_after = cpu_oracle.rhs(name, g["t"][k], g["y"][k], g["p"][k])
        assert np.array_equal(dy, g["dy"][k], equal_nan=True), (name, k)
        assert np.array_equal(p_after, g["p_after"][k], equal_nan=True), (name, k)


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_oracle_scheme_matches_reference_trajectory(name):
    g = np.load(os.path.join(GOLDEN, f"traj_{name}.npz"))
    S, P = g["states0"].copy(), g["params0"].copy()
    t, dt = 0.0, float(g["dt"])
    for _ in range(int(g["n_steps"])):
        assert cpu_oracle.step(name, S, P, t, dt, int(g["n_sub"]), 2) == 0
        t = t + dt
    assert np.array_equal(S, g["states"])
    assert np.array_equal(P, g["params"])


def test_oracle_threads_do_not_change_results():
    g = np.load(os.path.join(GOLDEN, "traj_hh_tissue.npz"))
    out = []
    for threads in (1, 3):
        S, P = np.tile(g["states0"], (5, 1)), np.tile(g["params0"], (5, 1))
        cpu_oracle.step("hh_tissue", S, P, 0.0, 0.1, 25, threads)
        out.append((S, P))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_oracle_reports_nonfinite_rows():
    ns, np_ = cpu_oracle.dims("hh_test")
    S = np.zeros((3, ns))
    S[:, 3] = [-70.0, np.nan, -60.0]
    P = np.tile([120.0, 36.0, 0.1, 0.4, 53.2, -93.4, 1.0, 0, 0, 0, 0, 3.32, 12.83, 2.0, 7.7, 50.0, 70.9], (3, 1))
    assert cpu_oracle.step("hh_test", S, P, 0.0, 0.1, 25, 1) == 1


def test_idealized_hh_fixed_point():
    """K1: the idealized HH initial state with run_2D.py:190-195 concentrations is a fixed point."""
    from workloads import SETUP, builtin
    ode = builtin("hh_ideal")
    p = ode.init_parameter_values()
    for k, v in {**SETUP["hh_ideal"]["uniform"], **SETUP["hh_ideal"]["varying"]}.items():
        p[ode.parameter_indices(k)] = v
    y = ode.init_state_values()
    dy, p_after = cpu_oracle.rhs("hh_ideal", 0.0, y, p)
    assert np.max(np.abs(dy)) < 1e-11
    assert abs(p_after[15] + p_after[16]) < 1e-13
    S, P = y[None, :].copy(), p[None, :].copy()
    t = 0.0
    for _ in range(100):
        cpu_oracle.step("hh_ideal", S, P, t, 1e-4, 25, 1)
        t += 1e-4
    assert np.max(np.abs(S[0] - y) / np.abs(y)) < 1e-10
