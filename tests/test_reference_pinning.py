"""Live pinning against the reference tree (build container only; skipped elsewhere).

* the builtin model modules and the C oracle equal the reference's ``rhs_numba``
  cfuncs bit for bit on random points;
* the code generator's DAG, read from the REFERENCE's own mm_*.py source files and
  evaluated with Python floats, equals the cfunc bit for bit -- i.e. the generator
  accepts the reference modules unchanged and understands them the way numba does.
"""
import ctypes

import numpy as np
import pytest

from conftest import MODEL_NAMES, load_reference_module

P = ctypes.POINTER(ctypes.c_double)


def _points(ref, name, n, seed):
    from workloads import SETUP
    rng = np.random.default_rng(seed)
    y0, p0 = ref.init_state_values(), ref.init_parameter_values()
    cfg = SETUP[name]
    for k, v in {**cfg["uniform"], **cfg["varying"]}.items():
        p0[ref.parameter_indices(k)] = v
    try:
        p0[ref.parameter_indices("stim_amplitude")] = 3.0
    except ValueError:
        pass
    Y = y0 * (1 + 0.3 * rng.uniform(-1, 1, (n, len(y0))))
    if len(y0) >= 4:
        Y[:, :3] = rng.uniform(0, 1, (n, 3))
    Pm = p0 * (1 + 0.1 * rng.uniform(-1, 1, (n, len(p0))))
    T = rng.uniform(0, 400 * cfg["dt"], n)
    return T, Y, Pm


def _cfunc_eval(mod, t, y, p):
    y, p = y.copy(), p.copy()
    dy = np.zeros(len(y))
    mod.rhs_numba.ctypes(t, y.ctypes.data_as(P), dy.ctypes.data_as(P), p.ctypes.data_as(P))
    return dy, p


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_builtin_and_oracle_equal_reference_cfunc(reference_root, name):
    from oracle import cpu_oracle
    from workloads import builtin
    ref = load_reference_module(reference_root, name)
    mine = builtin(name)
    assert np.array_equal(ref.init_state_values(), mine.init_state_values())
    assert np.array_equal(ref.init_parameter_values(), mine.init_parameter_values())
    T, Y, Pm = _points(ref, name, 2000, 11)
    for k in range(len(T)):
        want = _cfunc_eval(ref, T[k], Y[k], Pm[k])
        got_o = cpu_oracle.rhs(name, T[k], Y[k], Pm[k])
        got_m = _cfunc_eval(mine, T[k], Y[k], Pm[k])
        for got in (got_o, got_m):
            assert np.array_equal(got[0], want[0], equal_nan=True)
            assert np.array_equal(got[1], want[1], equal_nan=True)


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_generator_reads_reference_source(reference_root, name):
    import os
    from knpemi_b200.codegen import parse_model_source
    from knpemi_b200.codegen.interpret import evaluate
    from knpemi_b200.models import REFERENCE_FILE
    ref = load_reference_module(reference_root, name)
    path = os.path.join(reference_root, REFERENCE_FILE[name])
    pm = parse_model_source(open(path).read(), filename=path)
    T, Y, Pm = _points(ref, name, 500, 5)
    for k in range(len(T)):
        want_dy, want_p = _cfunc_eval(ref, T[k], Y[k], Pm[k])
        dy, p_after = evaluate(pm, T[k], Y[k], Pm[k])
        assert np.array_equal(np.array(dy), want_dy, equal_nan=True)
        assert np.array_equal(np.array(p_after), want_p, equal_nan=True)


def test_index_function_protocol_matches_reference(reference_root):
    from workloads import builtin
    for name in MODEL_NAMES:
        ref, mine = load_reference_module(reference_root, name), builtin(name)
        names_s = [n for n, _ in mine.STATES]
        names_p = [n for n, _ in mine.PARAMETERS]
        for n in names_s:
            assert ref.state_indices(n) == mine.state_indices(n)
        for n in names_p:
            assert ref.parameter_indices(n) == mine.parameter_indices(n)
        if len(names_s) > 1:
            assert ref.state_indices(*names_s[:2]) == mine.state_indices(*names_s[:2])
        with pytest.raises(ValueError):
            mine.parameter_indices("no_such_parameter")
        with pytest.raises(ValueError):
            ref.parameter_indices("no_such_parameter")
