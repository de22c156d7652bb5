"""The generator's emitted model code, compiled for the HOST, against the oracle (CPU only).

`tests/native/host_twin/kem_kernel.cuh` stands in for the step kernel's header, so g++ compiles
the very text `codegen/emit.py` writes for nvcc -- hoisted section, right-hand side with the
shared exponentials / `a*rcp(b)` / relaxed gates, outputs, time-only factors, literal table --
against the host build of `csrc/kem_math.cuh`.  The GPU parity tests (`test_gpu_parity.py`)
check the same code as device code; this module lets the CPU suite see a generator or
`kem_math` regression before any GPU does.  Same tolerance (1e-10) and floors as the GPU tests.

What the host build cannot show: the reciprocal / rsqrt seeds (MUFU on the device, a truncated
division here -- equally coarse, `kem_math.cuh:rcp_seed`) and nvcc's FMA contraction of the
emitted expressions (off here), both far below the tolerance."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from workloads import SETUP, builtin, synthetic_tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "knp-emi-fenics-x_b200", "csrc")
TWIN = os.path.join(ROOT, "tests", "native", "host_twin")
MODELS = ("hh_ideal", "hh_tissue", "glial_tissue", "glial_bench", "calibration", "hh_test")
RTOL, STATE_FLOOR, CURRENT_FLOOR = 1e-10, 1e-6, 3e-4         # tests/test_gpu_parity.py


def compile_twin(source, workdir, tag, contract="off"):
    src, lib = os.path.join(workdir, f"twin_{tag}.cpp"), os.path.join(workdir, f"twin_{tag}.so")
    with open(src, "w") as f:
        f.write(source)
    r = subprocess.run(["g++", "-O2", "-mfma", f"-ffp-contract={contract}", "-shared", "-fPIC", "-x", "c++",
                        "-I", TWIN, "-I", CSRC, "-o", lib, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(lib)
    L.twin_step_rk4.restype = C.c_int
    L.twin_step_rk4.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]
    L.twin_name.restype = C.c_char_p
    return L


def twin_step(L, S, P, t0, dt, n_sub, device_tonly=False):
    assert S.flags.c_contiguous and P.flags.c_contiguous
    return L.twin_step_rk4(len(S), S.ctypes.data, P.ctypes.data, t0, dt, n_sub, int(device_tonly))


def rel_err(got, want, frac):
    floor = np.maximum(frac * np.max(np.abs(want), axis=0, keepdims=True), 1e-300)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor)))


def run_pair(name, L, n, n_steps, n_sub=25, device_tonly=False):
    from oracle import cpu_oracle
    ode, cfg = builtin(name), SETUP[name]
    S, P, _, mask = synthetic_tables(name, n)
    S2, P2 = S.copy(), P.copy()
    c_stim = ode.parameter_indices("stim_amplitude")
    t = 0.0
    for _ in range(n_steps):
        P[mask, c_stim] = cfg["stim"]                           # sticky stimulus, odeSolver.py:110-112
        P2[mask, c_stim] = cfg["stim"]
        assert cpu_oracle.step(name, S, P, t, cfg["dt"], n_sub) == 0
        assert twin_step(L, S2, P2, t, cfg["dt"], n_sub, device_tonly) == 0
        t = t + cfg["dt"]
    return S2, P2, S, P


@pytest.fixture(scope="module")
def twins(tmp_path_factory):
    from knpemi_b200.codegen.build import generate
    work = str(tmp_path_factory.mktemp("host_twin"))
    cache = {}

    def get(name, **opts):
        from knpemi_b200.codegen import EmitOptions
        key = (name, tuple(sorted(opts.items())))
        if key not in cache:
            em = generate(builtin(name), EmitOptions(**opts) if opts else None)
            cache[key] = (compile_twin(em.source, work, f"{name}_{len(cache)}"), em)
        return cache[key]
    return get


@pytest.mark.parametrize("name", MODELS)
def test_emitted_code_matches_the_oracle_on_the_host(twins, name):
    """2 000 DOFs x 10 stimulated PDE steps, default options: the code the GPU runs."""
    L, em = twins(name)
    assert L.twin_name().decode() == name
    got_S, got_P, S, P = run_pair(name, L, 2000, 10)
    assert rel_err(got_S, S, STATE_FLOOR) < RTOL
    assert rel_err(got_P, P, CURRENT_FLOOR) < RTOL


@pytest.mark.parametrize("name", ("hh_ideal", "calibration"))
@pytest.mark.parametrize("opts", [dict(fuse_exp=False), dict(exact_div=True), dict(relax_gates=False),
                                  dict(fuse_exp=False, exact_div=True, relax_gates=False)],
                         ids=["exp_as_written", "exact_div", "gates_as_written", "all_rewrites_off"])
def test_every_rewrite_switch_keeps_the_tolerance(twins, name, opts):
    L, _ = twins(name, **opts)
    got_S, got_P, S, P = run_pair(name, L, 500, 10)
    assert rel_err(got_S, S, STATE_FLOOR) < RTOL
    assert rel_err(got_P, P, CURRENT_FLOOR) < RTOL


@pytest.mark.parametrize("name", ("hh_ideal", "hh_tissue", "calibration"))
def test_device_form_of_the_time_only_factors(twins, name):
    """Scheme O3 evaluates the stimulus envelope on the device (`tonly_dev`, kem::exp) because
    its stage times are not known on the host; the emitted device form must agree with the
    host form (libm) to rounding."""
    L, _ = twins(name)
    a = run_pair(name, L, 300, 10, device_tonly=False)
    b = run_pair(name, L, 300, 10, device_tonly=True)
    assert rel_err(b[0], a[0], STATE_FLOOR) < RTOL and rel_err(b[1], a[1], CURRENT_FLOOR) < RTOL


def test_fma_contraction_of_the_emitted_code_stays_within_the_tolerance(twins, tmp_path):
    """nvcc contracts a*b+c of the emitted expressions into FMAs; g++ with -ffp-contract=fast
    does the same on the host."""
    _, em = twins("hh_ideal")
    L = compile_twin(em.source, str(tmp_path), "contracted", contract="fast")
    got_S, got_P, S, P = run_pair("hh_ideal", L, 2000, 10)
    assert rel_err(got_S, S, STATE_FLOOR) < RTOL and rel_err(got_P, P, CURRENT_FLOOR) < RTOL


def test_golden_trajectories_on_the_host(twins):
    """The committed fixtures (the REFERENCE's cfuncs pushed through scheme O1)."""
    for name in MODELS:
        g = np.load(os.path.join(ROOT, "tests", "golden", f"traj_{name}.npz"))
        S, P = np.ascontiguousarray(g["states0"]).copy(), np.ascontiguousarray(g["params0"]).copy()
        L, _ = twins(name)
        t = 0.0
        for _ in range(int(g["n_steps"])):
            assert twin_step(L, S, P, t, float(g["dt"]), int(g["n_sub"])) == 0
            t = t + float(g["dt"])
        assert rel_err(S, g["states"], STATE_FLOOR) < RTOL, name
        assert rel_err(P, g["params"], CURRENT_FLOOR) < RTOL, name


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("mode", ["fast", "libm"])
def test_random_models_on_the_host(tmp_path, seed, mode):
    """The random straight-line models of `test_gpu_random_models.py` (arbitrary expressions of
    the supported language: every dependency class, integer powers, conditionals, extra libm
    calls), emitted and compiled for the host, against scheme O1 over the DAG interpreter."""
    from knpemi_b200.codegen import EmitOptions, parse_model_source
    from knpemi_b200.codegen.build import generate_from_source
    from test_gpu_random_models import NP, NS, numpy_rk4, random_model_source
    src = random_model_source(seed)
    pm = parse_model_source(src, filename=f"mm_random_{seed}.py")
    em = generate_from_source(src, f"mm_random_{seed}", NS, NP, EmitOptions(math=mode),
                              filename=f"mm_random_{seed}.py")
    L = compile_twin(em.source, str(tmp_path), f"random_{seed}_{mode}")
    n = 48
    rng = np.random.default_rng(seed)
    S = rng.uniform(-1, 1, (n, NS))
    P = np.tile(np.array([0.5, -0.3, 0.8, 0.0, 0.0]), (n, 1))
    P[:, 1] = rng.uniform(-1, 1, n)
    S2, P2 = S.copy(), P.copy()
    t = 0.5                                                       # the steps cross t = 0.7 and 0.93
    for _ in range(3):
        assert twin_step(L, S2, P2, t, 0.25, 7) == 0
        S, P = numpy_rk4(pm, S, P, t, 0.25, 7)
        t += 0.25
    assert np.all(np.isfinite(S))
    scale = lambda a: np.maximum(np.abs(a), 1e-3 * np.max(np.abs(a), axis=0, keepdims=True) + 1e-12)   # noqa: E731
    tol = 1e-10 if mode == "fast" else 1e-11
    assert np.max(np.abs(S2 - S) / scale(S)) < tol
    assert np.max(np.abs(P2 - P) / scale(P)) < tol
