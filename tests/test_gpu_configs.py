"""The remaining BASELINE.json configurations as parity / property tests.

configs[0]  2D idealized neuron, ~10^4 membrane DOFs (and the real 496-DOF mesh size)
configs[3]  calibration system to steady state (run_calibration.py:65-66: 10 000 steps of dt 0.1)
configs[4]  multi-tag tissue membrane: HH neurons (tag 1) + glia (tag 2), two models side by side
            (local_astrocyte_depolarization/run_stim_duration.py:171-181)
"""
import numpy as np
import pytest

from ducks_for_tests import Func, Space
from workloads import SETUP, builtin, load_tables, synthetic_tables

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def close(a, b, rtol=RTOL):
    scale = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b), axis=0, keepdims=True) + 1e-300)
    return float(np.max(np.abs(a - b) / scale)) < rtol


@pytest.mark.parametrize("n", [496, 10_000])
def test_config0_idealized_2d_hundred_steps(built, n):
    """100 PDE steps of dt = 1e-4 (run_2D.py:174-175) with the synaptic stimulus of
    run_2D.py:263-266, refreshed concentration traces every step."""
    from knpemi_b200.odeSolver import MembraneModel
    from oracle.membrane_oracle import OracleMembraneModel
    name = "hh_ideal"
    ode = builtin(name)
    S, P, X, mask = synthetic_tables(name, n, seed=42)
    gpu = MembraneModel(ode, None, 1, Space(X), verbose=False, devices=[0])
    cpu = OracleMembraneModel(ode, None, 1, Space(X), oracle_name=name)
    load_tables(gpu, S, P)
    cpu.states[:] = S
    cpu.parameters[:] = P
    rng = np.random.default_rng(0)
    loc = lambda x: x[0] < 20e-6     # noqa: E731
    for k in range(100):
        if k % 10 == 0:               # the PDE side slowly changes the traces
            trace = SETUP[name]["varying"]["K_e"] * (1 + 0.02 * rng.uniform(-1, 1, n))
            for m in (gpu, cpu):
                m.set_parameter("K_e", Func(trace))
        for m in (gpu, cpu):
            m.step_lsoda(1e-4, {'stim_amplitude': 10.0}, loc)
    assert close(np.asarray(gpu.states), cpu.states)
    assert close(np.asarray(gpu.parameters), cpu.parameters)
    assert gpu.time == cpu.time
    gpu.close()


def test_config3_calibration_runs_to_steady_state(built):
    """K3: from the embedded initial values the 14 states drift slowly
    (|RHS| <= 2.2e-2 on V_g) and settle; the GPU follows the oracle over the whole run."""
    from knpemi_b200.odeSolver import MembraneModel
    from oracle import cpu_oracle
    name = "calibration"
    ode = builtin(name)
    n_gpu, n_cpu, n_steps = 10_000, 64, 10_000
    gpu = MembraneModel(ode, None, 1, Space(np.zeros((n_gpu, 3))), verbose=False, devices=[0])
    S = np.tile(ode.init_state_values(), (n_cpu, 1))
    P = np.tile(ode.init_parameter_values(), (n_cpu, 1))
    t = 0.0
    for k in range(n_steps):
        gpu.step_async(0.1, {'stim_amplitude': 0})        # run_calibration.py:28-29,66
        if k % 500 == 499:
            gpu.synchronize()
    gpu.synchronize()
    for k in range(n_steps):
        cpu_oracle.step(name, S, P, t, 0.1, 25, 0)
        t += 0.1
    got = np.asarray(gpu.states)
    assert np.array_equal(got[0], got[-1])                 # identical DOFs stay identical
    assert close(got[:n_cpu], S, rtol=1e-9)                # 10^6 RK4 sub-steps of round-off
    # steady: one more step changes nothing beyond 1e-7 relative
    before = got[0].copy()
    gpu.step_lsoda(0.1, {'stim_amplitude': 0})
    after = np.asarray(gpu.states)[0]
    assert np.max(np.abs(after - before) / np.abs(before)) < 1e-7
    assert gpu.time == pytest.approx(1000.1)
    gpu.close()


def test_config4_two_tags_side_by_side(built):
    """Neuron membrane (tissue HH, tag 1) and glial membrane (mm_glial, tag 2) as two models
    in one process, stepped alternately like solve_odes does per tag (run_stim_duration.py:92-124);
    a third model on the same space (benchmark case, tags 5/6/7 on one Q) is independent."""
    from knpemi_b200.odeSolver import MembraneModel
    from oracle.membrane_oracle import OracleMembraneModel
    models = []
    for tag, name, n in ((1, "hh_tissue", 30_011), (2, "glial_tissue", 20_003), (5, "glial_tissue", 20_003)):
        S, P, X, mask = synthetic_tables(name, n, seed=tag)
        g = MembraneModel(builtin(name), None, tag, Space(X), verbose=False, devices=[0])
        c = OracleMembraneModel(builtin(name), None, tag, Space(X), oracle_name=name)
        load_tables(g, S, P)
        c.states[:] = S
        c.parameters[:] = P
        models.append((name, g, c))
    assert [g.tag for _, g, _ in models] == [1, 2, 5]
    for step in range(5):
        for name, g, c in models:
            stim = {'stim_amplitude': SETUP[name]["stim"]}
            for m in (g, c):
                m.step_lsoda(0.1, stim, lambda x: x[0] < 20e-6)
    for name, g, c in models:
        assert close(np.asarray(g.states), c.states)
        assert close(np.asarray(g.parameters), c.parameters)
        g.close()
