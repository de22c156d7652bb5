"""The reference's own glue code, executed verbatim (build container only).

``src/knpemi/odeSolver.py`` (the reference ``MembraneModel``), ``src/knpemi/utils.py``
(``setup_membrane_model`` :105-148, ``update_ode_variables`` :210-235) and the model module
``examples/idealized_geometries/mm_hh.py`` are imported from /root/reference through the stubs
of tests/shims/ and run by tests/golden/make_glue_transcript.py.  Here:

* the committed fixtures (call transcript + arrays) are what that run produces today;
* the restated class the GPU tests use as their checker (oracle/membrane_oracle.py) gives the
  same values when the reference's glue drives IT instead of the reference class.

The GPU replay of the transcript is tests/test_gpu_reference_glue.py.
"""
import functools
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module")
def glue(reference_root):
    pytest.importorskip("numba")
    import make_glue_transcript as g
    ode_solver, utils, mm_hh = g.load_reference()
    return g, ode_solver, utils, mm_hh


def test_committed_transcript_is_what_the_reference_glue_does(glue):
    g, ode_solver, utils, mm_hh = glue
    calls, arrays, S, P, X = g.run_glue(ode_solver.MembraneModel, mm_hh)
    with open(os.path.join(HERE, "golden", "glue_hh_ideal.json")) as f:
        meta = json.load(f)
    assert json.loads(json.dumps(calls)) == meta["calls"]
    fix = np.load(os.path.join(HERE, "golden", "glue_hh_ideal.npz"))
    assert set(arrays) | {"dof_coordinates", "final_states", "final_parameters"} == set(fix.files)
    for k, a in arrays.items():
        assert np.array_equal(a, fix[k]), k
    assert np.array_equal(S, fix["final_states"]) and np.array_equal(P, fix["final_parameters"])
    # the sequence solve_odes drives per PDE step (run_2D.py:88-109): 6 traces + phi_M in,
    # one step, phi_M + 3 currents out
    names = [c["method"] for c in meta["calls"]]
    assert names.count("step_lsoda") == 4 and names.count("set_parameter") == 24
    assert names.count("set_membrane_potential") == 3          # not at k = 0 (utils.py:233)
    assert names[:5] == ["set_parameter_values"] * 5           # Cm, psi, z_K, z_Cl, z_Na (utils.py:124-129)


def test_restated_class_equals_the_reference_class_under_the_reference_glue(glue):
    """oracle/membrane_oracle.py stands in for the reference class in every GPU parity test:
    driven by the reference's own utils.py it must reproduce the reference class's values."""
    from oracle.membrane_oracle import OracleMembraneModel
    g, ode_solver, utils, mm_hh = glue
    fix = np.load(os.path.join(HERE, "golden", "glue_hh_ideal.npz"))
    restated = functools.partial(OracleMembraneModel, oracle_name="hh_ideal")
    calls, arrays, S, P, X = g.run_glue(restated, mm_hh)
    for k, a in arrays.items():
        if k.startswith("in") and "set_membrane_potential" not in k:
            assert np.array_equal(a, fix[k]), k                 # same traces were offered
            # (phi_M handed back at k > 0 is the model's own output of the step before)
        else:
            scale = np.maximum(np.abs(fix[k]), 1e-6 * np.max(np.abs(fix[k])) + 1e-300)
            assert np.max(np.abs(a - fix[k]) / scale) < 1e-12, k
    assert np.allclose(S, fix["final_states"], rtol=1e-12, atol=0)
    # setters, getters and the sticky stimulus are plain copies: bitwise
    ode = mm_hh
    for name in ("K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i", "stim_amplitude", "Cm", "psi", "z_Na"):
        c = ode.parameter_indices(name)
        assert np.array_equal(P[:, c], fix["final_parameters"][:, c]), name
