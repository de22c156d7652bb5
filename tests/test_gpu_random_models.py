"""Randomised differential test of the whole product path: random straight-line model source
-> generator -> nvcc -> fused kernel -> C ABI, against an independent NumPy implementation of
scheme O1 driven by the generator's own DAG *interpreter* (Python floats, glibc libm).

The six shipped models exercise one family of expressions; this checks that the generator's
dependency classes (hoisted / host-evaluated / in-loop), constant tables, integer powers and
the fast-math rewrites are right for arbitrary expressions of the supported language."""
import importlib.util
import math
import random

import numpy as np
import pytest

from ducks_for_tests import Space

pytestmark = pytest.mark.gpu

NS, NP = 3, 5          # parameters[3], parameters[4] are output slots


def random_expr(rng, depth, leaves):
    if depth == 0 or rng.random() < 0.25:
        return rng.choice(leaves)
    kind = rng.random()
    a = random_expr(rng, depth - 1, leaves)
    b = random_expr(rng, depth - 1, leaves)
    if kind < 0.25:
        return f"({a} + {b})"
    if kind < 0.45:
        return f"({a} - {b})"
    if kind < 0.65:
        return f"({a} * {b})"
    if kind < 0.75:
        return f"({a} / (1.5 + ({b})**2))"                       # division with a safe divisor
    if kind < 0.85:
        return f"math.exp(-(({a})**2) / {rng.choice(['3.0', '7', '0.9'])})"   # bounded exp
    if kind < 0.90:
        return f"np.sqrt(1.0 + ({a})**2)"
    if kind < 0.94:
        return f"math.log(2.0 + ({a})**2)"
    if kind < 0.955:
        return f"(0.5 + ({a})**2) ** 1.5"
    if kind < 0.965:
        return f"(({a}) if ({b}) > 0.05 else ({b}) * 0.5)"                  # conditional expression
    if kind < 0.975:
        return f"math.tanh({a})"
    if kind < 0.985:
        return f"(abs({a}) + min({a}, {b}) - np.maximum({b}, 0.1))"
    return f"math.pow({a}, {rng.choice([2, 3, 4])})"


def random_model_source(seed):
    rng = random.Random(seed)
    consts = ["0.1", "2", "0.37", "1.0e-1", "3.25", "-0.7"]
    p_leaves = [f"parameters[{c}]" for c in range(3)] + consts
    s_leaves = [f"states[{c}]" for c in range(NS)] + p_leaves
    t_leaves = ["t", "0.3", "2.0"]
    lines = ["import math", "import numpy as np", "",
             "def rhs_numba(t, states, values, parameters):"]
    lines.append(f"    a = {random_expr(rng, 2, p_leaves)}")                          # parameter-only
    lines.append(f"    g = np.exp(-np.mod(t, 0.7) / 0.4) * (t < 0.93) + {random_expr(rng, 1, t_leaves)} * 0.01")
    lines.append(f"    b = {random_expr(rng, 4, s_leaves + ['a'])}")
    lines.append(f"    c = {random_expr(rng, 4, s_leaves + ['a', 'b', 'g'])}")
    lines.append("    parameters[3] = b - c")
    lines.append("    parameters[4] = a + g")
    for k in range(NS):                                         # keep the dynamics contracting
        e = random_expr(rng, 3, ['a', 'g'] + s_leaves)
        lines.append(f"    values[{k}] = -states[{k}] + 0.2 * np.exp(-(({e})**2)) * (b - {k + 1} * c) / (1.0 + b**2 + c**2)")
    src = "\n".join(lines) + "\n"
    src += (f"\ndef init_state_values(**kw):\n    return np.array([0.1, -0.2, 0.3])\n"
            f"def init_parameter_values(**kw):\n    return np.array([0.5, -0.3, 0.8, 0.0, 0.0])\n"
            "def state_indices(*n):\n    d = {'x': 0, 'y': 1, 'V': 2}\n    r = [d[k] for k in n]\n    return r if len(r) > 1 else r[0]\n"
            "def parameter_indices(*n):\n    d = {'p0': 0, 'p1': 1, 'p2': 2, 'o0': 3, 'o1': 4}\n"
            "    r = [d[k] for k in n]\n    return r if len(r) > 1 else r[0]\n")
    return src


def numpy_rk4(pm, S, P, t0, dt, n_sub):
    """Scheme O1 over the DAG interpreter, row by row (the independent implementation)."""
    from knpemi_b200.codegen.interpret import evaluate
    S, P = S.copy(), P.copy()
    h = dt / n_sub
    for r in range(len(S)):
        y, p = list(S[r]), list(P[r])
        f = lambda t, yy: np.array(evaluate(pm, t, yy, p)[0])     # noqa: E731
        y = np.array(y)
        for j in range(n_sub):
            ta, tb, tc = t0 + j * h, t0 + (j + 0.5) * h, t0 + (j + 1.0) * h
            k1 = f(ta, y)
            k2 = f(tb, y + 0.5 * h * k1)
            k3 = f(tb, y + 0.5 * h * k2)
            k4 = f(tc, y + h * k3)
            y = y + h / 6.0 * (((k1 + 2 * k2) + 2 * k3) + k4)
        _, p_after = evaluate(pm, t0 + dt, y, p)
        S[r], P[r] = y, p_after
    return S, P


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("mode", ["fast", "libm"])
def test_random_model_matches_interpreter(built, tmp_path, seed, mode):
    from knpemi_b200.codegen import EmitOptions, parse_model_source
    from knpemi_b200.odeSolver import MembraneModel
    src = random_model_source(seed)
    path = tmp_path / f"mm_random_{seed}.py"
    path.write_text(src)
    spec = importlib.util.spec_from_file_location(path.stem, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pm = parse_model_source(src, filename=str(path))
    n = 96
    rng = np.random.default_rng(seed)
    S = rng.uniform(-1, 1, (n, NS))
    P = np.tile(mod.init_parameter_values(), (n, 1))
    P[:, 1] = rng.uniform(-1, 1, n)                               # one per-DOF parameter
    m = MembraneModel(mod, None, 1, Space(np.zeros((n, 3))), verbose=False, devices=[0], n_sub=7,
                      emit_options=EmitOptions(math=mode))
    for c in range(NS):
        m.states[:, c] = S[:, c]
    m.parameters[:, 1] = P[:, 1]
    t = 0.5                                                       # the steps cross t = 0.7 and 0.93
    for _ in range(3):
        m.time = t
        m.step_lsoda(0.25, None)
        S, P = numpy_rk4(pm, S, P, t, 0.25, 7)
        t += 0.25
    got_S, got_P = np.asarray(m.states), np.asarray(m.parameters)
    m.close()
    assert np.all(np.isfinite(S))
    scale = lambda a: np.maximum(np.abs(a), 1e-3 * np.max(np.abs(a), axis=0, keepdims=True) + 1e-12)   # noqa: E731
    tol = 1e-10 if mode == "fast" else 1e-11
    assert np.max(np.abs(got_S - S) / scale(S)) < tol
    assert np.max(np.abs(got_P - P) / scale(P)) < tol
