"""Builtin model modules follow the plugin protocol of the reference's mm_*.py files
(init_* / *_indices behaviour: mm_hh.py:7-131)."""
import numpy as np
import pytest

from conftest import MODEL_NAMES
from workloads import builtin

DIMS = {"hh_ideal": (4, 22), "hh_tissue": (4, 22), "glial_tissue": (1, 23), "glial_bench": (1, 21),
        "calibration": (14, 13), "hh_test": (4, 17)}


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_tables_and_indices(name):
    ode = builtin(name)
    y, p = ode.init_state_values(), ode.init_parameter_values()
    assert (len(y), len(p)) == DIMS[name]
    assert y.dtype == np.float64 and p.dtype == np.float64
    # fresh arrays each call (the reference builds the tables row by row from them)
    y[0] = 123.0
    assert ode.init_state_values()[0] != 123.0
    for k, (nm, default) in enumerate(ode.STATES):
        assert ode.state_indices(nm) == k and ode.init_state_values()[k] == default
    for k, (nm, default) in enumerate(ode.PARAMETERS):
        assert ode.parameter_indices(nm) == k
    first_two = [nm for nm, _ in ode.PARAMETERS[:2]]
    assert ode.parameter_indices(*first_two) == [0, 1]           # list for several names


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_errors_and_overrides(name):
    ode = builtin(name)
    with pytest.raises(ValueError, match="Unknown state"):
        ode.state_indices("nope")
    with pytest.raises(ValueError, match="Unknown param"):
        ode.parameter_indices("nope")
    with pytest.raises(ValueError, match="is not a parameter"):
        ode.init_parameter_values(nope=1.0)
    with pytest.raises(ValueError, match="is not a state"):
        ode.init_state_values(nope=1.0)
    nm = ode.PARAMETERS[0][0]
    assert ode.init_parameter_values(**{nm: 42.0})[0] == 42.0


def test_v_index_convention():
    for name in MODEL_NAMES:
        ode = builtin(name)
        if name == "calibration":
            with pytest.raises(ValueError):
                ode.state_indices("V")          # V_n / V_g instead
        else:
            assert ode.state_indices("V") == len(ode.STATES) - 1


def test_rhs_numba_is_a_cfunc_with_the_lsoda_signature():
    import ctypes
    ode = builtin("hh_test")
    assert isinstance(ode.rhs_numba.address, int) and ode.rhs_numba.address != 0
    P = ctypes.POINTER(ctypes.c_double)
    y, p = ode.init_state_values(), ode.init_parameter_values()
    dy = np.zeros(4)
    ode.rhs_numba.ctypes(0.0, y.ctypes.data_as(P), dy.ctypes.data_as(P), p.ctypes.data_as(P))
    # SURVEY.md A.6 probe values
    assert dy[3] == pytest.approx(-0.422, abs=1e-3)
    assert p[8] == pytest.approx(1.4357, abs=1e-4) and p[9] == pytest.approx(-1.0138, abs=1e-4)
