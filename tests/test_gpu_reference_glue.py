"""Replay of the reference glue's call transcript on the CUDA backend (``-m gpu``).

tests/golden/glue_hh_ideal.{json,npz} were produced in the build container by running the
reference's own ``utils.setup_membrane_model`` / ``update_ode_variables`` / ``solve_odes``
sequence around the reference's own ``MembraneModel`` (tests/golden/make_glue_transcript.py).
Here every recorded call is made on ``knpemi_b200.odeSolver.MembraneModel`` -- through the
``knpemi.odeSolver`` import path a reference user would keep -- and every value the reference
read back is compared.
"""
import json
import os
import sys

import numpy as np
import pytest

from ducks_for_tests import Func, Space

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-10


def _err(got, want, floor_frac):
    scale = np.maximum(np.abs(want), floor_frac * np.max(np.abs(want)) + 1e-300)
    return float(np.max(np.abs(got - want) / scale))


@pytest.mark.parametrize("exchange", ["immediate", "deferred"])
def test_transcript_of_the_reference_glue_replays_on_the_gpu(built, exchange):
    root = os.path.dirname(HERE)
    compat = os.path.join(root, "knp-emi-fenics-x_b200", "compat")
    for k in [k for k in sys.modules if k == "knpemi" or k.startswith("knpemi.")]:
        del sys.modules[k]                                   # (a synthetic package of the CPU glue tests)
    sys.path.insert(0, compat)
    try:
        from knpemi.odeSolver import MembraneModel          # the reference's import path
    finally:
        sys.path.remove(compat)
    from knpemi_b200.models import hh_ideal
    with open(os.path.join(HERE, "golden", "glue_hh_ideal.json")) as f:
        meta = json.load(f)
    fix = np.load(os.path.join(HERE, "golden", "glue_hh_ideal.npz"))
    m = MembraneModel(hh_ideal, None, 1, Space(fix["dof_coordinates"]), verbose=False, devices=[0],
                      exchange=exchange)
    n = meta["n_dof"]
    checked = 0
    for call in meta["calls"]:
        name = call["method"]
        if name == "set_parameter_values":
            m.set_parameter_values({k: (lambda x, v=v: v) for k, v in call["values"].items()})
        elif name == "set_parameter":
            m.set_parameter(call["which"], Func(fix[call["array"]]))
        elif name == "set_membrane_potential":
            m.set_membrane_potential(Func(fix[call["array"]]))
        elif name == "step_lsoda":
            loc = eval("lambda x: " + call["locator"]) if call["locator"] else None
            m.step_lsoda(dt=call["dt"], stimulus=call["stimulus"], stimulus_locator=loc)
        elif name in ("get_parameter", "get_membrane_potential"):
            u = Func(np.zeros(n))
            if name == "get_parameter":
                assert m.get_parameter(call["which"], u) is u
            else:
                assert m.get_membrane_potential(u) is u
            want = fix[call["expect"]]
            floor = 1e-3 if name == "get_parameter" else 1e-6     # currents / potential (tests/diag_parity.py)
            assert _err(u.x.array, want, floor) < RTOL, (call, _err(u.x.array, want, floor))
            checked += 1
        else:
            raise AssertionError(f"unexpected call in the transcript: {name}")
    assert checked == 19
    S, P = np.asarray(m.states), np.asarray(m.parameters)
    for c in range(S.shape[1]):
        assert _err(S[:, c], fix["final_states"][:, c], 1e-6) < RTOL, c
    for nm in ("K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i", "stim_amplitude", "Cm", "psi", "z_Na", "z_K", "z_Cl"):
        c = hh_ideal.parameter_indices(nm)
        assert np.array_equal(P[:, c], fix["final_parameters"][:, c]), nm     # copies: bitwise
    assert m.time == pytest.approx(meta["n_steps"] * meta["dt"], rel=1e-15)
    m.close()
