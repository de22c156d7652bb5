"""RHS -> CUDA generator: parser, DAG, dependency classes, emission (CPU only)."""
import math
import os

import numpy as np
import pytest

from conftest import MODEL_NAMES
from knpemi_b200 import codegen
from knpemi_b200.codegen import EmitOptions, ModelSourceError, generate_from_source, parse_model_source
from knpemi_b200.codegen.interpret import evaluate
from workloads import builtin

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _src(body, header="def rhs_numba(t, states, values, parameters):\n"):
    return "import math\nimport numpy as np\n" + header + "".join("    " + ln + "\n" for ln in body)


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_dag_evaluates_like_the_reference_cfunc(name):
    """Parser + DAG semantics (folding, integer powers, association) against the golden
    RHS vectors that came from the reference's numba cfuncs -- bit for bit."""
    ode = builtin(name)
    pm = parse_model_source(open(ode.__file__).read(), filename=ode.__file__)
    g = np.load(os.path.join(GOLDEN, f"rhs_{name}.npz"))
    for k in range(len(g["t"])):
        dy, p_after = evaluate(pm, g["t"][k], g["y"][k], g["p"][k])
        assert np.array_equal(np.array(dy), g["dy"][k], equal_nan=True)
        assert np.array_equal(np.array(p_after), g["p_after"][k], equal_nan=True)


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_emitted_model_facts(name):
    ode = builtin(name)
    em = codegen.generate(ode)
    assert em.ns == len(ode.init_state_values()) and em.np == len(ode.init_parameter_values())
    expected_out = {"hh_ideal": [15, 16, 17], "hh_tissue": [15, 16, 17], "glial_tissue": [5, 6, 7],
                    "glial_bench": [9, 10, 11], "calibration": [], "hh_test": [8, 9, 10]}[name]
    assert em.out_cols == expected_out
    assert "kem_model_descriptor" not in em.source.split("KEM_DEFINE_MODEL")[0]
    assert em.source.count("KEM_DEFINE_MODEL(") == 1
    # time enters only through host-evaluated slots: no use of `t` in device code
    device = em.source.split("static void tonly")[0]
    assert "exp(" in device or name.startswith("glial") is False
    # glial models have no time dependence at all
    assert em.n_tslots == {"glial_tissue": 0, "glial_bench": 0, "calibration": 1}.get(name, 2)


def test_generation_is_deterministic_and_path_independent(tmp_path):
    ode = builtin("hh_tissue")
    a = codegen.generate(ode)
    b = codegen.generate(ode)
    assert a.source == b.source and a.source_hash == b.source_hash
    # same source under another directory gives the same translation unit
    src = open(ode.__file__).read()
    c = generate_from_source(src, "hh_tissue", a.ns, a.np, filename=os.path.basename(ode.__file__))
    assert c.source == a.source


def test_math_modes_emit_different_code():
    ode = builtin("hh_ideal")
    fast = codegen.generate(ode, EmitOptions(math="fast")).source
    libm = codegen.generate(ode, EmitOptions(math="libm")).source
    assert "kem::exp(" in fast and "kem::exp(" not in libm
    assert "kem::rcp(" in fast and "kem::div(" not in fast and " / " in libm
    exact = codegen.generate(ode, EmitOptions(math="fast", exact_div=True)).source
    assert "kem::div(" in exact
    # a*(1 - x) - b*x is emitted as a - x*(a + b) unless asked not to
    plain = codegen.generate(ode, EmitOptions(math="fast", relax_gates=False))
    assert len(codegen.generate(ode).stats["relaxed_gates"]) == 3 and not plain.stats["relaxed_gates"]
    assert plain.stats["deriv"]["sub"] > codegen.generate(ode).stats["deriv"]["sub"]
    with pytest.raises(ValueError):
        codegen.generate(ode, EmitOptions(math="wrong"))


def test_dependency_classes_and_hoisting():
    body = [
        "a = parameters[0] * parameters[1]",              # parameter-only -> hoisted
        "g = np.exp(-np.mod(t, 2.0)) * (t < 5)",           # time-only      -> host slot
        "c = math.exp(1.0) + 2",                          # constant       -> folded
        "values[0] = a * states[0] + g * parameters[2] + c",
        "parameters[3] = a - states[0]",
    ]
    em = generate_from_source(_src(body), "toy", 1, 4)
    assert em.out_cols == [3] and em.used_cols == [0, 1, 2]
    assert em.n_tslots == 1
    hoist = em.source.split("void hoist")[1].split("void deriv")[0]
    deriv = em.source.split("void deriv")[1].split("void outputs")[0]
    assert "p[0] * p[1]" in hoist and "p[0]" not in deriv
    assert "exp" not in deriv                               # both exps live elsewhere
    import math
    assert float.hex(math.exp(1.0) + 2) in em.source        # folded constant, exact literal ...
    assert "KC[" in deriv                                   # ... read from the constant table
    tonly = em.source.split("static void tonly")[1]
    assert "kem_npmod_host(t" in tonly and "exp(" in tonly


def test_integer_power_lowering_matches_numba_order():
    pm = parse_model_source(_src(["values[0] = states[0]**3 + math.pow(states[0], 4) + states[0]**1.5"]))
    x = 1.2345678901234567
    dy, _ = evaluate(pm, 0.0, [x], [])
    x2 = x * x
    import math
    assert dy[0] == (x * x2 + x2 * x2) + math.pow(x, 1.5)
    em = generate_from_source(_src(["values[0] = states[0]**3"]), "p3", 1, 0)
    assert "y[0] * y[0]" in em.source and "pow(" not in em.source.split("tonly")[0]


def test_docstrings_and_comment_blocks_are_skipped():
    body = ['"""docstring"""', "x = states[0]", '"""', "old = code(here)", '"""', "values[0] = -x"]
    em = generate_from_source(_src(body), "doc", 1, 0)
    assert em.ns == 1


@pytest.mark.parametrize("body,msg", [
    (["for i in range(3):", "    values[0] = 1.0"], "only simple assignments"),
    (["values[0] = math.gamma(states[0])"], "unsupported call"),
    (["values[0] = helper(states[0])"], "unsupported call"),
    (["values[0] = states[i]"], "integer literals"),
    (["values[0] = undefined_name"], "not defined"),
    (["values[0] = [states[0]][0]"], "unsupported subscript"),
    (["values[0] = (lambda x: x)(states[0])"], "unsupported call"),
    (["values[0] = 0 < states[0] < 1"], "only single"),
    (["parameters[0] = 1.0", "values[0] = parameters[0]"], "read after"),
    (["x = 1.0"], "assigns no values"),
    (["values[1] = 1.0"], "outside the 1 states"),
    (["values[0] = 1.0/0.0"], "constant expression"),
])
def test_unsupported_constructs_are_rejected(body, msg):
    with pytest.raises(ModelSourceError, match=msg):
        generate_from_source(_src(body), "bad", 1, 1)


def test_missing_rhs_function():
    with pytest.raises(ModelSourceError, match="no top-level function"):
        parse_model_source("def other(t, y, dy, p):\n    dy[0] = 0.0\n")


def test_model_without_source_file_is_refused():
    class Fake:
        __name__ = "fake"

        @staticmethod
        def init_state_values():
            return np.zeros(1)

        @staticmethod
        def init_parameter_values():
            return np.zeros(1)
    with pytest.raises(ModelSourceError, match="__file__"):
        codegen.generate(Fake)


@pytest.mark.parametrize("name", ["hh_ideal", "calibration"])
def test_committed_sample_of_generated_code_is_current(name):
    """docs/generated_<model>.cu is what the generator emits today (kept for readers who want
    to see the device code without running anything)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "docs", f"generated_{name}.cu")) as f:
        assert f.read() == codegen.generate(builtin(name)).source


def test_conditionals_and_extra_libm_calls():
    """Beyond what the reference's six models use, but common in Gotran-generated modules:
    `a if c else b`, and/or, ==/!=, np.where, abs/tanh/min/max/... (CUDA libm on the device)."""
    import math
    body = [
        "v = states[0]",
        "gate = 1.0 if v > 0.5 else 0.25",
        "both = (v > 0.1 and parameters[0] != 0) or v == 7",
        "w = np.where(v < 0, -v, v) + abs(v - 1) + math.tanh(v) + min(v, 0.3) + np.maximum(v, 0.6)",
        "values[0] = gate * w + both + math.floor(2.5) + np.log1p(v * v) + math.atan2(v, 2.0)",
        "parameters[1] = math.fmod(v, 0.3) + math.cos(0.0)",
    ]
    src = _src(body)
    pm = parse_model_source(src)
    for v, p0 in ((0.7, 1.0), (0.2, 0.0), (-0.4, 2.0), (7.0, 0.0)):
        dy, p_after = evaluate(pm, 0.0, [v], [p0, 0.0])
        gate = 1.0 if v > 0.5 else 0.25
        both = float((v > 0.1 and p0 != 0) or v == 7)
        w = (-v if v < 0 else v) + abs(v - 1) + math.tanh(v) + min(v, 0.3) + max(v, 0.6)
        want = gate * w + both + 2.0 + math.log1p(v * v) + math.atan2(v, 2.0)
        assert dy[0] == pytest.approx(want, rel=1e-15)
        assert p_after[1] == pytest.approx(math.fmod(v, 0.3) + 1.0, rel=1e-15)
    em = generate_from_source(src, "cond", 1, 2)
    dev = em.source.split("static void tonly")[0]
    for token in ("tanh(", "fabs(", "fmin(", "fmax(", "log1p(", "atan2(", "fmod(", "!= 0.0 ?"):
        assert token in dev, token
    assert "floor(" not in dev                     # folded constant


def test_gotran_boilerplate_is_accepted():
    """Unmodified Gotran output wraps the expressions in argument guards and tuple unpacking;
    the reference authors stripped those for numba, other model files may still carry them."""
    body = [
        '"""Compute the right hand side"""',
        "import math",
        "assert(len(states) == 2)",
        "m, V = states",
        "assert(len(parameters) == 3)",
        "g, E, I_out = parameters",
        "if values is None:",
        "    values = np.zeros((2,), dtype=np.float64)",
        "else:",
        "    assert isinstance(values, np.ndarray) and values.shape == (2,)",
        "i = g * m * (V - E)",
        "i += 0.5",
        "parameters[2] = i",
        "values[0] = (1 - m) * math.exp(-V / 10) - m",
        "values[1] = -i",
        "return values",
    ]
    pm = parse_model_source(_src(body))
    dy, p_after = evaluate(pm, 0.0, [0.25, -60.0], [2.0, -80.0, 0.0])
    import math
    cur = 2.0 * 0.25 * (-60.0 + 80.0) + 0.5
    assert dy == [(1 - 0.25) * math.exp(60.0 / 10) - 0.25, -cur] and p_after[2] == cur
    em = generate_from_source(_src(body), "gotran", 2, 3)
    assert em.out_cols == [2] and em.used_cols == [0, 1]


# ------------------------------------------------------------------ shared exponentials
@pytest.mark.parametrize("name", MODEL_NAMES)
def test_shared_exponential_rewrite_stays_within_its_error_bound(name):
    """codegen/fuse_exp.py: the rewritten right-hand side, evaluated with Python floats, against
    the DAG as written, on the golden RHS inputs plus random physiological ones: every rewritten
    exponential within 2e-14 relative, and no amplified error in the derivatives."""
    from knpemi_b200.codegen.fuse_exp import fuse_exponentials
    ode = builtin(name)
    pm = parse_model_source(open(ode.__file__).read(), filename=ode.__file__)
    fused, report = fuse_exponentials(pm)
    if name.startswith("glial"):
        assert report == [] and fused is pm          # two exps of incommensurate slopes: untouched
        return
    # HH rates: exp((25-u)/10) - 1 stays as written, the two other /10 exponentials are shifts of
    # it, exp(-u/18), exp(-u/20), exp(-u/80) are powers 40, 36, 9 of exp(-u/720)
    assert [r["kind"] for r in report] == ["shift", "shift", "chain"]
    assert report[2]["powers"] == [9, 36, 40] and sum(r["exps_replaced"] for r in report) == 5
    g = np.load(os.path.join(GOLDEN, f"rhs_{name}.npz"))
    rng = np.random.default_rng(11)
    cases = [(g["t"][k], g["y"][k], g["p"][k]) for k in range(len(g["t"]))]
    for k in range(200):
        y = np.array(g["y"][k % len(g["t"])])
        y *= 1.0 + 0.3 * rng.uniform(-1, 1, y.shape)
        cases.append((g["t"][0], y, g["p"][k % len(g["t"])]))
    from knpemi_b200.codegen.parse import ParsedModel
    pairs = [pr for r in report for pr in r["nodes"]]
    as_written = ParsedModel(pm.dag, {k: a for k, (a, _) in enumerate(pairs)}, {}, "", 0)
    rewritten = ParsedModel(pm.dag, {k: b for k, (_, b) in enumerate(pairs)}, {}, "", 0)
    worst_rate = worst_dy = 0.0
    for t, y, p in cases:
        a, pa = evaluate(pm, t, y, p)
        b, pb = evaluate(fused, t, y, p)
        a, b = np.array(a), np.array(b)
        if not np.all(np.isfinite(a)):
            continue
        ea, _ = evaluate(as_written, t, y, p)
        eb, _ = evaluate(rewritten, t, y, p)
        worst_rate = max(worst_rate, float(np.max(np.abs(np.array(ea) - eb) / np.abs(ea))))
        # the derivatives: differences of rate terms, and x/(exp(x) - 1) amplifies by 1/|x|
        scale = np.maximum(np.abs(a), 1e-3 * np.max(np.abs(a)) + 1e-300)
        worst_dy = max(worst_dy, float(np.max(np.abs(a - b) / scale)))
        assert np.array_equal(np.array(pa), np.array(pb), equal_nan=True)       # currents: no exp
    print(name, "worst rate error", worst_rate, "worst derivative error", worst_dy)
    assert worst_rate < 2e-14, worst_rate
    assert worst_dy < 1e-11, worst_dy


def test_shared_exponential_groups_chain_and_limits():
    from knpemi_b200.codegen.fuse_exp import _chain, _commensurate, fuse_exponentials
    steps = _chain([9, 36, 40, 72])
    have = {1}
    for k, a, b in steps:
        assert a in have and b in have and a + b == k
        have.add(k)
    assert {9, 36, 40, 72} <= have and len(steps) == 8
    assert _commensurate([-100.0, -1000 / 18, -50.0, -12.5], 96)[1] == [72, 40, 36, 9]
    assert _commensurate([1.0, -2.0], 96) is None                 # opposite signs
    assert _commensurate([1.0, math.pi], 96) is None              # incommensurate
    assert _commensurate([1.0, 97.0], 96) is None                 # power above the limit
    # offsets that depend on parameters are hoisted, not left in the loop
    src = _src(["a = math.exp((states[0] - parameters[0]) / 10.0)",
                "b = math.exp((states[0] + 3.0) / 5.0)",
                "c = math.exp(states[0] / 2.5)",
                "values[0] = a * b * c"])
    pm = parse_model_source(src)
    fused, report = fuse_exponentials(pm)
    assert report[0]["powers"] == [1, 2] and report[0]["exps_replaced"] == 2      # constants only
    fused, report = fuse_exponentials(pm, param_offsets=True)
    assert report[0]["powers"] == [1, 2, 4] and report[0]["exps_replaced"] == 3
    em = generate_from_source(src, "fuse_probe", 1, 1, EmitOptions(fuse_exp_param_offsets=True))
    loop = em.source[em.source.index("void deriv"):em.source.index("void outputs")]
    assert loop.count("kem::exp") == 1
    hoist = em.source[em.source.index("void hoist"):em.source.index("void deriv")]
    assert hoist.count("kem::exp") == 1
    for y0 in (-3.0, 0.1, 7.5):
        a, _ = evaluate(pm, 0.0, [y0], [1.25])
        b, _ = evaluate(fused, 0.0, [y0], [1.25])
        assert abs(a[0] - b[0]) <= 1e-14 * abs(a[0])
    # off: every exp as written
    em = generate_from_source(src, "fuse_probe", 1, 1, EmitOptions(fuse_exp=False))
    assert em.source[em.source.index("void deriv"):em.source.index("void outputs")].count("kem::exp") == 3
    em = generate_from_source(src, "fuse_probe", 1, 1, EmitOptions(math="libm"))
    assert em.source[em.source.index("void deriv"):em.source.index("void outputs")].count("exp(") == 3


def test_shared_exponential_rewrite_on_random_rate_expressions():
    """Random right-hand sides built from exponentials of affine functions of two states, with
    rational slopes, constant and parameter-dependent offsets and every kind of consumer: each
    rewritten exponential stays within the documented bound of the one as written, and
    exponentials that feed a difference are never derived from a chain."""
    from knpemi_b200.codegen.fuse_exp import fuse_exponentials
    from knpemi_b200.codegen.parse import ParsedModel
    rng = np.random.default_rng(2024)
    n_fused = n_models = 0
    for trial in range(int(os.environ.get("KNPEMI_FUSE_TRIALS", "60"))):
        body, terms = [], []
        base = [float(rng.choice([0.5, 1.0, 2.5, 12.5, 40.0])) for _ in range(2)]
        for k in range(int(rng.integers(2, 8))):
            s = int(rng.integers(0, 2))
            num, den = int(rng.integers(1, 9)), int(rng.choice([1, 1, 2, 3, 4, 9]))
            sign = "-" if rng.random() < 0.7 else ""
            off = rng.choice(["", " + 1.5", " - 0.25", " + parameters[0]", " - parameters[1] / 3.0"])
            form = rng.integers(0, 3)
            arg = {0: f"({sign}states[{s}]{off}) * {num}.0 / {den * base[s]}",
                   1: f"({sign}{num}.0 * states[{s}] / {den}.0{off}) / {base[s]}",
                   2: f"{sign}states[{s}] / {den * base[s] / num}{off}"}[int(form)]
            body.append(f"e{k} = math.exp({arg})")
            use = rng.integers(0, 5)
            terms.append({0: f"3.0 * e{k}", 1: f"(e{k} - 1.0)", 2: f"1.0 / (e{k} + 1.0)",
                          3: f"states[{s}] / (1.0 - e{k})", 4: f"e{k} / (2.0 + states[{1 - s}])"}[int(use)])
        body.append("values[0] = " + " + ".join(terms[::2]))
        body.append("values[1] = " + (" * ".join(terms[1::2]) if terms[1::2] else "states[0]"))
        pm = parse_model_source(_src(body))
        fused, report = fuse_exponentials(pm, param_offsets=bool(trial % 2))
        n_models += 1
        if not report:
            continue
        n_fused += 1
        parents = {}
        for nid in pm.dag.reachable(list(pm.dy.values())):
            for c in pm.dag.nodes[nid].args:
                parents.setdefault(c, []).append(pm.dag.nodes[nid].op)
        for r in report:
            if r["kind"] == "chain":
                assert max(r["powers"]) <= 96
                for a, _ in r["nodes"]:
                    assert "sub" not in parents.get(a, []), "an exp feeding a difference was chained"
        pairs = [pr for r in report for pr in r["nodes"]]
        written = ParsedModel(pm.dag, {k: a for k, (a, _) in enumerate(pairs)}, {}, "", 0)
        rewritten = ParsedModel(pm.dag, {k: b for k, (_, b) in enumerate(pairs)}, {}, "", 0)
        for _ in range(20):
            y, p = rng.uniform(-3.0, 3.0, 2), rng.uniform(-1.0, 1.0, 2)
            ea, _ = evaluate(written, 0.0, y, p)
            eb, _ = evaluate(rewritten, 0.0, y, p)
            err = np.abs(np.array(ea) - eb) / np.abs(ea)
            assert np.all(err < 2 * 96 * 1.2e-16), (trial, err.max(), body)
            da, _ = evaluate(pm, 0.0, y, p)
            db, _ = evaluate(fused, 0.0, y, p)
            assert np.all(np.isfinite(db) == np.isfinite(da))
    assert n_fused >= n_models // 3, (n_fused, n_models)


# ------------------------------------------------------------------ affine collapse (experimental)
@pytest.mark.parametrize("name", MODEL_NAMES)
def test_affine_collapse_keeps_the_right_hand_side(name):
    """codegen/affine.py: chains of affine operations on one node become one FMA of that node.
    Evaluated with Python floats (mul and add rounded separately, i.e. without the FMA's extra
    accuracy) the rewritten right-hand side agrees with the one as written to rounding level."""
    from knpemi_b200.codegen.affine import collapse_affine
    from knpemi_b200.codegen.fuse_exp import fuse_exponentials
    ode = builtin(name)
    pm = parse_model_source(open(ode.__file__).read(), filename=ode.__file__)
    g = np.load(os.path.join(GOLDEN, f"rhs_{name}.npz"))
    rng = np.random.default_rng(5)
    for start in (pm, fuse_exponentials(pm)[0]):
        col, report = collapse_affine(start)
        assert report, "every builtin model rescales its membrane potential"
        assert all(r["operations_replaced"] >= 2 for r in report)
        worst = 0.0
        for k in range(150):
            y = np.array(g["y"][k % len(g["t"])])
            p = g["p"][k % len(g["t"])]
            if k >= len(g["t"]):
                y *= 1.0 + 0.2 * rng.uniform(-1, 1, y.shape)
            a, pa = evaluate(start, g["t"][0], y, p)
            b, pb = evaluate(col, g["t"][0], y, p)
            a, b, pa, pb = np.array(a), np.array(b), np.array(pa), np.array(pb)
            if not np.all(np.isfinite(a)):
                continue
            scale = np.maximum(np.abs(a), 1e-3 * np.max(np.abs(a)) + 1e-300)
            worst = max(worst, float(np.max(np.abs(a - b) / scale)))
            pscale = np.maximum(np.abs(pa), 1e-3 * np.max(np.abs(pa)) + 1e-300)
            assert np.all(np.abs(pa - pb) / pscale < 1e-12)
        assert worst < 1e-11, worst


def test_affine_collapse_is_off_by_default_and_changes_the_source_when_on():
    ode = builtin("hh_ideal")
    off = codegen.generate(ode)
    on = codegen.generate(ode, EmitOptions(collapse_affine=True))
    assert off.stats["collapsed_affine"] == [] and "collapse_affine" not in off.source
    assert on.stats["collapsed_affine"] and "collapse_affine=1" in on.source
    assert on.source_hash != off.source_hash
    loop = on.source[on.source.index("void deriv"):on.source.index("void outputs")]
    assert "// u" not in loop            # the rescaled potential u = 1e3*(V + 65e-3) is gone


def test_relaxed_gate_form_is_the_same_function():
    """codegen/relax.py: `a*(1 - x) - b*x -> a - x*(a + b)` on the DAG, checked with the
    interpreter on random points (equal up to the rounding of three operations)."""
    from knpemi_b200.codegen.interpret import evaluate
    from knpemi_b200.codegen.relax import relax_gates
    body = [
        "a = np.exp(states[1])",
        "b = parameters[0] * states[1] + 2.0",
        "values[0] = a * (1 - states[0]) - b * states[0]",
        "values[1] = (1.0 - states[1]) * b - states[1] * a",       # factors in the other order
        "parameters[1] = a * (1 - states[0]) - b * states[1]",      # not the pattern: different x
    ]
    pm = parse_model_source(_src(body))
    pm2, report = relax_gates(pm)
    assert len(report) == 2
    rng = np.random.default_rng(0)
    for _ in range(200):
        y = list(rng.uniform(0.01, 0.99, 2))
        p = [float(rng.uniform(0.5, 2.0)), 0.0]
        dy1, p1 = evaluate(pm, 0.0, y, list(p))
        dy2, p2 = evaluate(pm2, 0.0, y, list(p))
        assert np.allclose(dy1, dy2, rtol=1e-15, atol=1e-15) and p1[1] == p2[1]


@pytest.mark.parametrize("seed", range(40))
def test_random_model_source_parses_to_what_python_computes(seed, tmp_path):
    """Parser and DAG against CPython itself: the random straight-line models of the GPU
    differential test (`test_gpu_random_models.py`, which compares the kernel with the DAG
    interpreter) are executed as ordinary Python and compared with the interpreter, so a parse
    error (precedence, sign, a wrong dependency class, a mis-lowered call) cannot hide on both
    sides of that test.  Not bit-for-bit: the interpreter follows numba (`x**3` by squaring),
    CPython calls libm `pow`."""
    import importlib.util
    from test_gpu_random_models import NP, NS, random_model_source
    from knpemi_b200.codegen import parse_model_source
    from knpemi_b200.codegen.interpret import evaluate
    src = random_model_source(seed)
    path = tmp_path / f"mm_random_{seed}.py"
    path.write_text(src)
    spec = importlib.util.spec_from_file_location(path.stem, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pm = parse_model_source(src, filename=str(path))
    rng = np.random.default_rng(1000 + seed)
    worst = 0.0
    for t in (0.0, 0.31, 0.69, 0.71, 0.92, 0.94, 1.5):             # both sides of mod(t, 0.7) and t < 0.93
        y = rng.uniform(-1, 1, NS)
        p = mod.init_parameter_values()
        p[:3] = rng.uniform(-1, 1, 3)
        p_py, dy_py = p.copy(), np.zeros(NS)
        mod.rhs_numba(t, y.copy(), dy_py, p_py)
        dy, p_after = evaluate(pm, t, list(y), list(p))
        got, want = np.array(list(dy) + list(p_after)), np.concatenate([dy_py, p_py])
        assert len(p_after) == NP
        worst = max(worst, float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3))))
    assert worst < 1e-12, worst
