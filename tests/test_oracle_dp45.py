"""CPU twin of scheme O3 (oracle/knpemi_oracle.c:dp45_row): accuracy, economy, failure."""
import numpy as np
import pytest

from oracle import cpu_oracle
from workloads import SETUP, builtin, synthetic_tables


@pytest.mark.parametrize("name", ["hh_ideal", "hh_tissue", "glial_bench", "calibration"])
def test_dp45_meets_its_tolerance_with_fewer_rhs_calls(name):
    n, n_steps = 300, 5
    S, P, X, mask = synthetic_tables(name, n, seed=8)
    P[mask, builtin(name).parameter_indices("stim_amplitude")] = SETUP[name]["stim"]
    S1, P1, S2, P2, hs = S.copy(), P.copy(), S.copy(), P.copy(), np.zeros(n)
    t, dt, acc, rej = 0.0, SETUP[name]["dt"], 0, 0
    for _ in range(n_steps):
        bad, a, r = cpu_oracle.step_dp45(name, S1, P1, hs, t, dt, 1e-8, 1e-10, 2)
        assert bad == 0
        acc, rej = acc + a, rej + r
        cpu_oracle.step(name, S2, P2, t, dt, 800, 2)
        t += dt
    scale = np.maximum(np.abs(S2), 1e-3 * np.abs(S2).max(axis=0))
    assert (np.abs(S1 - S2) / scale).max() < 1e-7
    assert 6 * (acc + rej) / (n * n_steps) + 1 < 80          # scheme O1 spends 101
    assert np.all(hs > 0) and np.all(hs <= dt)
    # the output slots hold the currents at the end state (last call is the FSAL evaluation)
    out = [c for c, (nm, _) in enumerate(builtin(name).PARAMETERS) if nm.startswith("I_ch_")]
    if out:
        y, p = S1[0].copy(), P1[0].copy()
        _, p_after = cpu_oracle.rhs(name, t, y, p)
        assert np.allclose(p_after[out], P1[0][out], rtol=1e-12, atol=1e-300)


def test_dp45_threads_do_not_change_results():
    S, P, X, mask = synthetic_tables("hh_tissue", 500, seed=2)
    res = []
    for th in (1, 3):
        s, p, hs = S.copy(), P.copy(), np.zeros(500)
        cpu_oracle.step_dp45("hh_tissue", s, p, hs, 0.0, 0.1, 1e-8, 1e-10, th)
        res.append((s, p, hs))
    assert all(np.array_equal(a, b) for a, b in zip(res[0], res[1]))


def test_dp45_reports_failed_rows():
    S, P, X, mask = synthetic_tables("hh_test", 4, seed=2)
    S[2, 3] = np.nan
    hs = np.zeros(4)
    bad, acc, rej = cpu_oracle.step_dp45("hh_test", S, P, hs, 0.0, 0.1)
    assert bad == 1 and hs[2] == 0.0
