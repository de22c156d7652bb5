"""Device build of kem_math.cuh against CUDA's own libm / IEEE division, through a model
whose 'dynamics' are just exp and divisions of the state (so the comparison runs through
the real generator + kernel + C ABI path in both math modes)."""
import importlib.util
import textwrap

import numpy as np
import pytest

from ducks_for_tests import Func, Space

pytestmark = pytest.mark.gpu

SRC = textwrap.dedent('''
    import math
    import numpy as np
    def init_state_values(**values):
        return np.array([0.0, 1.0], dtype=np.float64)
    def init_parameter_values(**values):
        return np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0], dtype=np.float64)
    def state_indices(*names):
        d = {"V": 0, "w": 1}
        r = [d[n] for n in names]
        return r if len(r) > 1 else r[0]
    def parameter_indices(*names):
        d = {"o_exp": 0, "o_div": 1, "o_rcp": 2, "o_sing": 3, "o_log": 4, "o_sqrt": 5, "o_p15": 6}
        r = [d[n] for n in names]
        return r if len(r) > 1 else r[0]
    def rhs_numba(t, states, values, parameters):
        parameters[0] = math.exp(states[0])
        parameters[1] = states[0] / states[1]
        parameters[2] = 1.0 / states[1]
        parameters[3] = states[0] / (math.exp(states[0]) - 1.0)
        parameters[4] = math.log(states[1] * states[1])
        parameters[5] = np.sqrt(states[1] * states[1])
        parameters[6] = (states[1] * states[1]) ** 1.5
        values[0] = 0.0 * states[0]
        values[1] = 0.0 * states[1]
''')


def _outputs(tmp_path, math, x, w, **opts):
    from knpemi_b200.codegen import EmitOptions
    from knpemi_b200.odeSolver import MembraneModel
    path = tmp_path / "mm_math_probe.py"
    path.write_text(SRC)
    spec = importlib.util.spec_from_file_location("mm_math_probe", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    m = MembraneModel(mod, None, 1, Space(np.zeros((len(x), 3))), verbose=False, devices=[0],
                      emit_options=EmitOptions(math=math, **opts), n_sub=1)
    m.set_state('V', Func(x))
    m.set_state('w', Func(w))
    m.step_lsoda(1.0, None)
    out = np.asarray(m.parameters)
    m.close()
    return out


def test_device_exp_div_rcp_within_one_ulp_of_libm(built, tmp_path):
    rng = np.random.default_rng(0)
    n = 400_000
    x = np.concatenate([rng.uniform(-700, 700, n // 2), rng.uniform(-30, 30, n // 4),
                        rng.uniform(-1e-3, 1e-3, n // 4)])
    w = np.ldexp(rng.uniform(1, 2, n), rng.integers(-200, 200, n)) * rng.choice([-1.0, 1.0], n)
    fast = _outputs(tmp_path, "fast", x, w)
    exact = _outputs(tmp_path, "fast", x, w, exact_div=True)
    libm = _outputs(tmp_path, "libm", x, w)
    with np.errstate(over="ignore"):
        want_exp = np.exp(x.astype(np.longdouble))
        want_div = x.astype(np.longdouble) / w.astype(np.longdouble)
        want_rcp = 1.0 / w.astype(np.longdouble)
    # exp: table entry 0.5 ulp + final rounding 0.5 ulp + polynomial 0.09 ulp; a/b = a*rcp(b):
    # 0.51 ulp of the reciprocal + 0.5 ulp of the product, relative to a quotient up to 2x smaller
    for col, want, bound in ((0, want_exp, 1.1), (1, want_div, 1.6), (2, want_rcp, 1.0)):
        ulp = np.spacing(np.abs(want.astype(np.float64)))
        err_fast = np.abs(fast[:, col].astype(np.longdouble) - want) / ulp
        err_libm = np.abs(libm[:, col].astype(np.longdouble) - want) / ulp
        assert err_fast.max() <= bound, (col, float(err_fast.max()))
        assert err_libm.max() <= 1.0, (col, float(err_libm.max()))
    w2 = (w * w).astype(np.longdouble)       # w*w is what the device squared, rounded to double
    for col, want, bound in ((4, np.log(w2), 1.7), (5, np.sqrt(w2), 0.5 + 1e-9), (6, w2 * np.sqrt(w2), 1.3)):
        ulp = np.spacing(np.abs(want.astype(np.float64)))
        err_fast = np.abs(fast[:, col].astype(np.longdouble) - want) / ulp
        assert err_fast.max() <= bound, (col, float(err_fast.max()))
        err_libm = np.abs(libm[:, col].astype(np.longdouble) - want) / ulp
        assert err_libm.max() <= 2.0, (col, float(err_libm.max()))
    # EmitOptions(exact_div=True): residual-corrected division, correctly rounded like IEEE ->
    # identical bits; the one-step cubic reciprocal is faithful (<= 0.51 ulp): it may differ
    # from IEEE in the last bit, rarely
    assert np.array_equal(exact[:, 1], libm[:, 1])
    assert np.mean(fast[:, 2] != libm[:, 2]) < 0.02


def test_extra_libm_functions_and_conditionals_on_device(built, tmp_path):
    """Every function of codegen.ir.CALL1 / CALL2 plus select/and/or through the real path,
    against the DAG interpreter (Python's math module)."""
    from knpemi_b200.codegen import parse_model_source
    from knpemi_b200.codegen.interpret import evaluate
    from knpemi_b200.codegen.ir import CALL1, CALL2
    from knpemi_b200.odeSolver import MembraneModel
    one = sorted(CALL1)
    two = sorted(CALL2)
    n_out = len(one) + len(two) + 3
    lines = ["import math", "import numpy as np", "def init_state_values(**values):",
             "    return np.array([0.3, 0.7], dtype=np.float64)", "def init_parameter_values(**values):",
             f"    return np.zeros({n_out}, dtype=np.float64)", "def state_indices(*names):",
             "    d = {'V': 0, 'w': 1}", "    r = [d[n] for n in names]", "    return r if len(r) > 1 else r[0]",
             "def parameter_indices(*names):", "    return 0",
             "def rhs_numba(t, states, values, parameters):", "    x = states[0]", "    y = states[1]"]
    for k, f in enumerate(one):
        arg = {"asin": "x * 0.9", "acos": "x * 0.9", "log1p": "x * x", "log10": "x * x + 0.1"}.get(f, "x")
        lines.append(f"    parameters[{k}] = math.{f}({arg})")
    for k, f in enumerate(two):
        lines.append(f"    parameters[{len(one) + k}] = math.{f}(x, y)")
    base = len(one) + len(two)
    lines += [f"    parameters[{base}] = x if x > y else y * 2",
              f"    parameters[{base + 1}] = 1.0 * (x > 0 and y > 0.5) + 2.0 * (x < -0.5 or y == 0.25)",
              f"    parameters[{base + 2}] = np.where(x != y, x - y, 7.0)",
              "    values[0] = 0.0 * x", "    values[1] = 0.0 * y"]
    src = "\n".join(lines) + "\n"
    path = tmp_path / "mm_libm_probe.py"
    path.write_text(src)
    spec = importlib.util.spec_from_file_location("mm_libm_probe", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(5)
    n = 2000
    x = rng.uniform(-1, 1, n)
    y = rng.uniform(0.1, 1, n)
    y[:10] = 0.25
    x[10:20] = y[10:20]
    m = MembraneModel(mod, None, 1, Space(np.zeros((n, 3))), verbose=False, devices=[0], n_sub=1)
    m.set_state('V', Func(x))
    m.set_state('w', Func(y))
    m.step_lsoda(1.0, None)
    got = np.asarray(m.parameters)
    m.close()
    pm = parse_model_source(src)
    for r in range(0, n, 7):
        _, want = evaluate(pm, 1.0, [x[r], y[r]], [0.0] * n_out)
        want = np.array(want)
        assert np.allclose(got[r], want, rtol=1e-14, atol=1e-300), (r, np.argmax(np.abs(got[r] - want)))
