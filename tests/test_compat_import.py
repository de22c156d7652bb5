"""The reference's import path resolves to the B200 class (no device needed to import)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_path():
    code = ("import sys; sys.path[:0] = [%r, %r]; "
            "from knpemi.odeSolver import MembraneModel; import knpemi_b200.odeSolver as o; "
            "assert MembraneModel is o.MembraneModel; "
            "import inspect; sig = inspect.signature(MembraneModel.__init__); "
            "assert list(sig.parameters)[:5] == ['self', 'ode', 'ft', 'tag', 'Q']; "
            "sig = inspect.signature(MembraneModel.step_lsoda); "
            "assert list(sig.parameters) == ['self', 'dt', 'stimulus', 'stimulus_locator']; print('OK')"
            % (os.path.join(ROOT, "knp-emi-fenics-x_b200", "compat"), os.path.join(ROOT, "knp-emi-fenics-x_b200")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr


def test_public_method_signatures_match_the_reference_class():
    """Names, argument names and defaults of odeSolver.py:52-127."""
    import inspect
    from knpemi_b200.odeSolver import MembraneModel
    expect = {
        "set_state": ["self", "which", "u", "locator"], "set_parameter": ["self", "which", "u", "locator"],
        "get_state": ["self", "which", "u", "locator"], "get_parameter": ["self", "which", "u", "locator"],
        "set_state_values": ["self", "value_dict", "locator"],
        "set_parameter_values": ["self", "value_dict", "locator"],
        "set_membrane_potential": ["self", "u", "locator"], "get_membrane_potential": ["self", "u", "locator"],
        "step_lsoda": ["self", "dt", "stimulus", "stimulus_locator"],
    }
    for name, params in expect.items():
        sig = inspect.signature(getattr(MembraneModel, name))
        assert list(sig.parameters) == params, name
        last = list(sig.parameters.values())[-1]
        assert last.default is None, name
    assert isinstance(MembraneModel.V_index, property)
